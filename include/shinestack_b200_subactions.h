/*
 * shinestack_b200_subactions.h -- C ABI of the per-frame sub-actions BalanceFrames and Vignetting in
 * libshinestack_b200.so ("elementwise passes fused into the upload path").
 *
 * In the reference both are SubActions of CombinedActions (src/shinestacker/algorithms/stack_framework.py:246-302):
 * every frame is read, handed to each sub-action's run_frame(idx, ref_idx, img) and written back.  What is
 * data-parallel in them is one elementwise pass each:
 *   BalanceFrames -> CorrectionMapBase.adjust / apply_lut (balance.py:27-48): a 256- or 65536-entry look-up table
 *                    applied to every sample (one table for all channels, or one per channel);
 *   Vignetting    -> correct_vignetting (vignetting.py:71-97): every pixel divided by a radial sigmoid model.
 * The fits that produce the table / the model parameters (histograms, scipy bisect / curve_fit; balance.py:54-121,
 * vignetting.py:16-69) are scalar host work and stay on the host (shinestacker_b200/balance.py, vignetting.py).
 * A frame is uploaded once, the passes run on it in HBM, and it is either downloaded (sub-action protocol: the
 * reference writes the frame to a file) or handed on by device pointer (ss_push_frame_device), never leaving HBM.
 *
 * Conventions as in shinestack_b200.h.  Frames are H x W x 3 interleaved BGR, 8 or 16 bits per sample.
 */
#ifndef SHINESTACK_B200_SUBACTIONS_H
#define SHINESTACK_B200_SUBACTIONS_H

#include "shinestack_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sssub_handle sssub_handle;

int sssub_create(int32_t width, int32_t height, int32_t bytes_per_sample, int32_t device, sssub_handle **out);
void sssub_destroy(sssub_handle *h);

/* Host frame -> the handle's device frame (asynchronous; the host memory must stay valid until the next
 * sssub_download / sssub_sync). */
int sssub_upload(sssub_handle *h, const void *bgr, size_t row_pitch_bytes);
/* Replaces CorrectionMapBase.adjust (balance.py:34-48) on the device frame.  lut: host array of n_lut tables of
 * num_values (256 / 65536) samples of the frame's type; n_lut == 1: the one table maps every channel (LumiCorrection,
 * cv2.LUT / np.take on the whole image, balance.py:35-36), n_lut == 3: table c maps channel c in B, G, R order
 * (RGBCorrection, balance.py:43-47). */
int sssub_apply_lut(sssub_handle *h, const void *lut, int n_lut);
/* Replaces correct_vignetting (vignetting.py:71-97) on the device frame, all in float64 like numpy:
 * vignette = clip(sigmoid_model(r, i0, k, r0) / v0, 1e-6, 1), r = distance from (w/2, h/2); blended with
 * max_correction when it is < 1; 1 where min(B, G, R) < threshold; sample = clip(sample / vignette, 0, max),
 * truncated.  params = {i0, k, r0}; threshold is in sample units (black_threshold, x 256 for 16-bit frames). */
int sssub_apply_vignette(sssub_handle *h, const double *params, double v0, double max_correction, double threshold);
/* Device frame -> host (waits for everything queued on the handle). */
int sssub_download(sssub_handle *h, void *bgr, size_t row_pitch_bytes);
int sssub_sync(sssub_handle *h);
/* Device address / row pitch of the frame, for ss_push_frame_device / ssdm_push_frame_device / ssal_warp_device
 * (call sssub_sync first, or order the consumer after the handle's work yourself). */
int sssub_device_frame(sssub_handle *h, const void **d_bgr, size_t *pitch_bytes);
int sssub_kernel_time_ms(sssub_handle *h, float *ms);
int sssub_launch_count(sssub_handle *h, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* SHINESTACK_B200_SUBACTIONS_H */
