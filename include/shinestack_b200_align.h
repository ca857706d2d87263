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
/* Device address and row pitch of the last aligned frame (valid until the next ssal_warp on this handle). */
int ssal_result_device(ssal_handle *h, const void **d_bgr, size_t *pitch_bytes);
/* Device time of the last warp (CUDA events around its kernels) and kernels launched so far. */
int ssal_kernel_time_ms(ssal_handle *h, float *ms);
int ssal_launch_count(ssal_handle *h, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* SHINESTACK_B200_ALIGN_H */
