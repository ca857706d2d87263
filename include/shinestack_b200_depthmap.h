/*
 * shinestack_b200_depthmap.h -- C ABI of the depth-map stack algorithm in libshinestack_b200.so.
 *
 * Second stack algorithm behind the same boundary as shinestack_b200.h:
 * the reference's DepthMapStack (src/shinestacker/algorithms/depth_map.py:10-123) is, like PyramidStack,
 * a stack_algo object that FocusStack / FocusStackBunch drive through .focus_stack(filenames)
 * (reference stack.py:23,29,47,108).  The reference-side binding is the ctypes shim in
 * shinestacker_b200/depth_map.py.
 *
 * Conventions as in shinestack_b200.h: plain pointers and sizes, 0 on success or a negative ss_status,
 * ss_last_error() for the message, one thread at a time per handle, frames are H x W x 3 interleaved BGR
 * of 8 or 16 bits per sample.  No CPU fallback: ssdm_create fails with SS_ERR_NO_DEVICE without an
 * sm_100 GPU.
 */
#ifndef SHINESTACK_B200_DEPTHMAP_H
#define SHINESTACK_B200_DEPTHMAP_H

#include "shinestack_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssdm_handle ssdm_handle;

/* Constructor arguments of DepthMapStack (reference depth_map.py:11-26; defaults constants.py:144-157)
 * + the frame geometry the reference takes from the first file (depth_map.py:70). */
typedef struct ssdm_config {
    int32_t width, height;
    int32_t bytes_per_sample;  /* 1 (uint8) or 2 (uint16)                                              */
    int32_t map_type;          /* 0 = 'average' (default), 1 = 'max' (softmax with temperature)         */
    int32_t energy;            /* 0 = 'laplacian' (default), 1 = 'sobel'                                */
    int32_t kernel_size;       /* cv2.Laplacian ksize (default 5; odd, 5..31)                           */
    int32_t blur_size;         /* cv2.GaussianBlur ksize before the Laplacian (default 5; odd, 1..31)   */
    int32_t smooth_size;       /* cv2.bilateralFilter diameter (default 15; <= 0 disables smoothing)    */
    double temperature;        /* softmax temperature of map_type 'max' (default 0.1)                   */
    int32_t levels;            /* blend pyramid levels (default 3)                                      */
    int32_t float_bits;        /* 32 ('float-32'); 64 is rejected with SS_ERR_INVALID_OPTION            */
    int32_t device;            /* CUDA device ordinal                                                   */
    int32_t max_frames;        /* capacity of the device-resident stack                                 */
    int32_t reserved[4];
} ssdm_config;

/* Replaces DepthMapStack.__init__ + the first-file set-up of focus_stack (depth_map.py:11-26,64-71). */
int ssdm_create(const ssdm_config *cfg, ssdm_handle **out);
void ssdm_destroy(ssdm_handle *h);
/* Forget all frames so that the handle can fuse another stack of the same geometry (FocusStackBunch
 * reuses one stack_algo, reference stack.py:94-97). */
int ssdm_reset(ssdm_handle *h);

/* Replaces one iteration of the first loop of focus_stack (depth_map.py:68-77: read_img + img_bw) and that
 * frame's share of get_laplacian_map / get_sobel_map (depth_map.py:28-42): asynchronous upload into the
 * device-resident stack, then gray + focus energy + per-frame min/max.  The host buffer must stay valid
 * until ssdm_sync_copies / ssdm_finish_begin. */
int ssdm_push_frame(ssdm_handle *h, const void *bgr, size_t row_pitch_bytes);
int ssdm_push_frame_device(ssdm_handle *h, const void *d_bgr, size_t row_pitch_bytes);
int ssdm_sync_copies(ssdm_handle *h);

/* Replaces depth_map.py:88-93: normalisation by the global maximum energy, smooth_energy (bilateral
 * filter), get_focus_map. */
int ssdm_finish_begin(ssdm_handle *h);
/* Replaces one iteration of the second loop (depth_map.py:94-115): Gaussian pyramids of frame `frame` and
 * of its focus map, weighted Laplacian levels added to the blend.  Frames 0..n-1, once each, in order
 * (the host polls check_running between calls like the reference does between files). */
int ssdm_blend_frame(ssdm_handle *h, int frame);
/* Replaces depth_map.py:116-123: collapse, clip(|x|), astype; out_bgr (host, may be NULL) receives the
 * fused H x W x 3 image in the input sample type. */
int ssdm_finish_end(ssdm_handle *h, void *out_bgr, size_t out_row_pitch_bytes);
/* Convenience: finish_begin + blend_frame(0..n-1) + finish_end. */
int ssdm_finish(ssdm_handle *h, void *out_bgr, size_t out_row_pitch_bytes);

/* ---- Frames of one stack sharded over the GPUs of a box: one process and one handle per GPU, every rank pushes
 * its own contiguous run of the frames (ranks in stack order).  The three reductions over the stack of
 * depth_map.py:88-115 are split: the maximum energy crosses the ranks as one float through the host; the per-pixel sum
 * of the smoothed energies (get_focus_map 'average', :55-57) and the blended pyramid (:104-111) are read out of the
 * peers' device memory (CUDA IPC over NVLink); the pyramid is reduced in bands of rows, one per rank, which the root
 * then collects.  Sums run rank by rank, i.e. in stack order but associated differently
 * from one long chain (float32 rounding; the stacker's contract is 1 LSB).  map_type 'max' (softmax, :58-60) crosses
 * the ranks twice: the per-pixel maximum (exact in any order), then the sum of the exponentials.
 *
 *   ssdm_push_frame ... ; ssdm_shard_local_max  -> all-gather, max      -> ssdm_shard_finish_local(global max)
 *   ssdm_shard_export -> all-gather -> ssdm_shard_connect (once per handle)
 *   ssdm_sync; barrier; [map_type max: ssdm_shard_softmax_rel; ssdm_sync; barrier;]
 *   ssdm_shard_finish_weights; ssdm_blend_frame(0..n_local-1); ssdm_sync; barrier
 *   ssdm_shard_reduce_band; ssdm_sync; barrier
 *   root: ssdm_shard_finish_end(out); others: ssdm_shard_finish_end(NULL); barrier before the next stack.          */
int ssdm_sync(ssdm_handle *h);
int ssdm_shard_local_max(ssdm_handle *h, float *max_out);
int ssdm_shard_finish_local(ssdm_handle *h, float global_max);
int ssdm_shard_buffer_count(ssdm_handle *h);                      /* 2 + levels */
int ssdm_shard_export(ssdm_handle *h, void *blob /* count x 64 bytes */);
int ssdm_shard_connect(ssdm_handle *h, const void *blobs /* world x count x 64 */, int world, int rank, int root);
int ssdm_shard_ptrs(ssdm_handle *h, void **ptrs /* count */);     /* same-process variant of export / connect */
int ssdm_shard_connect_ptrs(ssdm_handle *h, void *const *ptrs /* world x count */, int world, int rank, int root);
int ssdm_shard_softmax_rel(ssdm_handle *h);                        /* no-op for map_type average */
int ssdm_shard_finish_weights(ssdm_handle *h);
int ssdm_shard_reduce_band(ssdm_handle *h);
int ssdm_shard_finish_end(ssdm_handle *h, void *out_bgr, size_t out_row_pitch_bytes);

/* Re-run the whole chain over the frames still resident in HBM, no download (bench.py). */
int ssdm_rerun_resident(ssdm_handle *h);
/* Test hook, H x W float32 of one frame: what = 2 raw energy (before finish_begin; get_laplacian_map /
 * get_sobel_map), 0 normalised energy (after finish_begin, when smoothing is on), 1 focus map
 * (get_focus_map, after finish_begin). */
int ssdm_get_plane(ssdm_handle *h, int what, int frame, float *out);
/* Device time (CUDA events on the compute stream) from the first push to the end of finish_end, and the
 * number of kernels launched since the last reset. */
int ssdm_kernel_time_ms(ssdm_handle *h, float *ms);
int ssdm_launch_count(ssdm_handle *h, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* SHINESTACK_B200_DEPTHMAP_H */
