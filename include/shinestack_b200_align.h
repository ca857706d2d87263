/*
 * shinestack_b200_align.h -- C ABI of the warp stage of AlignFrames in libshinestack_b200.so
 * (SURVEY.md section 8f, rank 2).
 *
 * The reference's AlignFrames (src/shinestacker/algorithms/align.py:254-353) is a per-frame sub-action:
 * run_frame(idx, ref_idx, img) -> align_images(img_ref, img) (align.py:170-252).  Feature detection,
 * matching and the RANSAC fit (align.py:118-168, cv2 SIFT / FLANN / estimateAffinePartial2D /
 * findHomography) stay on the host; this entry point replaces what follows once the transform is known
 * (align.py:228-251): cv2.warpAffine / cv2.warpPerspective of the frame, the validity mask, the 21x21
 * Gaussian blur and the paste of blurred pixels into the invalid border.  The reference-side binding is the
 * ctypes shim in shinestacker_b200/align.py.
 *
 * Conventions as in shinestack_b200.h.  Frames are H x W x 3 interleaved BGR, 8 or 16 bits per sample;
 * the destination has the size of the source, as in the reference (align.py:231,239: dsize = (w, h)).
 */
#ifndef SHINESTACK_B200_ALIGN_H
#define SHINESTACK_B200_ALIGN_H

#include "shinestack_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssal_handle ssal_handle;

enum { SSAL_RIGID = 0 /* constants.ALIGN_RIGID: 2x3 matrix, cv2.warpAffine */,
       SSAL_HOMOGRAPHY = 1 /* constants.ALIGN_HOMOGRAPHY: 3x3 matrix, cv2.warpPerspective */ };
enum { SSAL_BORDER_CONSTANT = 0, SSAL_BORDER_REPLICATE = 1, SSAL_BORDER_REPLICATE_BLUR = 2 };

int ssal_create(int32_t width, int32_t height, int32_t bytes_per_sample, int32_t device, ssal_handle **out);
void ssal_destroy(ssal_handle *h);

/* Replaces align.py:228-251 for one frame.  m: the forward transform exactly as find_transform returns it
 * (6 doubles row-major for SSAL_RIGID, 9 for SSAL_HOMOGRAPHY; the library inverts it like cv2 does);
 * border_value: alignment_config['border_value'] (4 doubles, may be NULL = 0); blur_kernel / blur_taps: the
 * integer kernel of cv2.GaussianBlur((21, 21), sigmaX=border_blur) for this sample type (8 fractional bits
 * for uint8, 16 for uint16, sums to 1.0), only read for SSAL_BORDER_REPLICATE_BLUR.  src_bgr and dst_bgr are
 * host memory; dst_bgr may be NULL, the aligned frame then stays in device memory (ssal_result_device) and
 * can be handed to ss_push_frame_device / ssdm_push_frame_device without leaving HBM. */
int ssal_warp(ssal_handle *h, const void *src_bgr, size_t src_pitch_bytes, const double *m, int transform,
              int border_mode, const double *border_value, const int32_t *blur_kernel, int blur_taps,
              void *dst_bgr, size_t dst_pitch_bytes);
/* Same with the source frame already in device memory. */
int ssal_warp_device(ssal_handle *h, const void *d_src_bgr, size_t src_pitch_bytes, const double *m, int transform,
                     int border_mode, const double *border_value, const int32_t *blur_kernel, int blur_taps,
                     void *dst_bgr, size_t dst_pitch_bytes);
/* ---- gradient-based estimator (BASELINE.json north-star: "the ECC/feature alignment pre-pass becomes a CUDA
 * warp-reduction over image gradients").  Optional replacement of the host-side feature pipeline
 * detect_and_compute + find_transform (align.py:89-148) for constants.ALIGN_RIGID: enhanced-correlation-
 * coefficient iterations on 8-bit gray pyramids, one reduction pass over the image gradients per iteration.
 * The reference has no such estimator; its accuracy contract is the reference's own precision test
 * (tests/test_0031_align_precision.py:62-65). */
/* The frame the following frames are aligned to (AlignFrames.run_frame: process.img_ref(ref_idx)). */
int ssal_ecc_set_reference(ssal_handle *h, const void *ref_bgr, size_t pitch_bytes);
/* Estimates the forward 2x3 similarity m_out (6 doubles; moving frame -> reference frame, what
 * cv2.estimateAffinePartial2D returns in the reference) of mov_bgr.  finest_level: pyramid level at which the
 * refinement stops (0 = full resolution, 1 = half, like the reference's default subsample = 2; < 0 = default 1);
 * max_iters per level (<= 0: 30); eps_px: stop when a step moves no pixel by more than this (<= 0: 1e-3);
 * init_m: optional initial forward transform.  info (4 doubles, may be NULL): correlation coefficient, total
 * iterations, overlap fraction, last step in pixels.  The uploaded frame stays on the device: a following
 * ssal_warp with src_bgr == NULL warps it without a second upload. */
int ssal_ecc_estimate(ssal_handle *h, const void *mov_bgr, size_t pitch_bytes, int finest_level, int max_iters,
                      double eps_px, const double *init_m, double *m_out, double *info);

/* Pipelining of a frame sequence (the reference processes the frames of AlignFrames one by one,
 * stack_framework.py:246-302; here the upload of frame i+1 hides behind the fit and warp of frame i):
 * ssal_prefetch starts the asynchronous upload of the NEXT frame (page-locked host memory) into a second source
 * buffer; ssal_use_prefetched makes it the current frame, which ssal_ecc_estimate / ssal_warp then take with a NULL
 * source pointer.  At most one prefetched frame can be outstanding. */
int ssal_prefetch(ssal_handle *h, const void *next_bgr, size_t pitch_bytes);
int ssal_use_prefetched(ssal_handle *h);
/* Waits for everything queued on this handle (uploads included): the host buffers handed to ssal_prefetch may then go. */
int ssal_sync(ssal_handle *h);

/* Device address and row pitch of the last aligned frame (outputs are double-buffered: valid until the SECOND next ssal_warp on this handle). */
int ssal_result_device(ssal_handle *h, const void **d_bgr, size_t *pitch_bytes);
/* Device time of the last warp (CUDA events around its kernels) and kernels launched so far. */
int ssal_kernel_time_ms(ssal_handle *h, float *ms);
int ssal_launch_count(ssal_handle *h, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* SHINESTACK_B200_ALIGN_H */
