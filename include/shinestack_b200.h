/*
 * shinestack_b200.h -- C ABI of libshinestack_b200.so: a B200-native (sm_100a CUDA) implementation of
 * the Laplacian-pyramid focus-stacking hot path of lucalista/shinestacker.
 *
 * The reference is pure Python; its "FFI" for this path is the stack-algo protocol that
 * FocusStack / FocusStackBunch call (reference src/shinestacker/algorithms/stack.py:23,29,47,108 and
 * base_stack_algo.py:9-42).  Each entry point below names the reference interface it replaces.  The
 * reference-side binding is the ctypes shim in shinestacker_b200/pyramid.py.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * ss_status; nothing throws across the ABI; ss_last_error() returns a thread-local message.  Frames are
 * H x W x 3 interleaved BGR, 8 or 16 bits per sample, exactly what cv2.imread delivers to
 * PyramidStack.focus_stack (reference pyramid.py:158,172).  All calls on one handle must come from one
 * thread at a time (the reference host calls its stack_algo synchronously, stack.py:29).
 */
#ifndef SHINESTACK_B200_H
#define SHINESTACK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ss_handle ss_handle;

typedef enum ss_status {
    SS_OK = 0,
    SS_ERR_INVALID_OPTION = -1, /* reference: InvalidOptionError (base_stack_algo.py:18-22)          */
    SS_ERR_SHAPE = -2,          /* reference: ShapeError (utils.py:60-61)                            */
    SS_ERR_BIT_DEPTH = -3,      /* reference: BitDepthError (utils.py:62-63)                         */
    SS_ERR_CUDA = -4,           /* CUDA runtime / driver failure (no reference equivalent)           */
    SS_ERR_NO_DEVICE = -5,      /* no sm_100 device: the product never falls back to a CPU path      */
    SS_ERR_STATE = -6,          /* call sequence violated (e.g. finish before any frame)             */
    SS_ERR_NOMEM = -7,
    SS_ERR_COMM = -8            /* multi-GPU exchange failure                                        */
} ss_status;

/* Constructor arguments of PyramidStack (reference pyramid.py:114-123) + the frame geometry that the
 * reference derives from the first file (pyramid.py:158-165) + placement. */
typedef struct ss_config {
    int32_t width;             /* W, pixels                                                         */
    int32_t height;            /* H, pixels                                                         */
    int32_t bytes_per_sample;  /* 1 (uint8) or 2 (uint16)                                           */
    int32_t min_size;          /* PyramidStack(min_size=32): levels = int(log2(min(H,W)/min_size))  */
    int32_t kernel_size;       /* PyramidStack(kernel_size=5): base-level entropy/variance window   */
    double gen_kernel;         /* PyramidStack(gen_kernel=0.4): Burt-Adelson 'a'                    */
    int32_t float_bits;        /* PyramidStack(float_type): 32 ('float-32') or 64 ('float-64')      */
    int32_t device;            /* CUDA device ordinal                                               */
    int32_t max_frames;        /* capacity of the device-resident frame store (>= frames pushed)    */
    int32_t batch_frames;      /* frames fused per kernel-chain launch (0 = default)                */
    int32_t frame_index_base;  /* global index of this handle's first frame (multi-GPU sharding)    */
    int32_t resident_frames;   /* device frame slots (0 = 2*batch_frames ring; >= frames pushed keeps
                                  the whole stack in HBM so that ss_rerun_resident can be used)       */
    int32_t reserved[4];
} ss_config;

/* Library / build identification. */
const char *ss_version(void);
/* Thread-local description of the last error on this thread. */
const char *ss_last_error(void);
/* Number of CUDA devices with compute capability 10.x visible to the process (may be 0). */
int ss_device_count(void);

/* Replaces PyramidStack.__init__ + the first-file set-up in focus_stack (pyramid.py:114-123,158-165).
 * Fails with SS_ERR_NO_DEVICE when no sm_100 GPU is present. */
int ss_create(const ss_config *cfg, ss_handle **out);
void ss_destroy(ss_handle *h);
/* Forget all frames / running state so the same handle can fuse another stack of the same geometry
 * (FocusStackBunch reuses one stack_algo for every bunch, stack.py:94-97). */
int ss_reset(ss_handle *h);

/* Geometry queries: number of Laplacian levels L actually built (pyramid.py:127-131,165) and the size
 * of level l (0..L; L is the base level). */
int ss_num_levels(const ss_handle *h);
int ss_level_size(const ss_handle *h, int level, int32_t *height, int32_t *width);

/* Pinned host staging memory for callers that want zero-copy-speed uploads (optional). */
int ss_host_alloc(void **ptr, size_t bytes);
int ss_host_free(void *ptr);
/* 1 when ptr lies in page-locked host memory known to CUDA (uploads from it are asynchronous), else 0. */
int ss_host_is_pinned(const void *ptr);

/* Replaces one iteration of the second loop of focus_stack (pyramid.py:170-177: read_img +
 * process_single_image): hands one host frame to the device.  The copy is asynchronous; the frame
 * memory must stay valid until the next ss_sync/ss_finish_* call.  Frames must be pushed in stack
 * order; when batch_frames frames are pending their kernel chain (pyramid build + energy + running
 * arg-max select, pyramid.py:125-139 and :48-55) is launched behind the copy. */
int ss_push_frame(ss_handle *h, const void *bgr, size_t row_pitch_bytes);
/* Same for a host frame whose samples are in R, G, B order -- the payload of an uncompressed RGB TIFF, which is what
 * the reference's write_img produces (utils.py:23-30): uploaded as it is, channels swapped on the device (the result is
 * the B, G, R frame cv2.imread would have delivered, utils.py:11-20), so the host side only moves bytes. */
int ss_push_frame_rgb(ss_handle *h, const void *rgb, size_t row_pitch_bytes);
/* Same, for a frame that already lives in device memory (device pointer, same layout). */
int ss_push_frame_device(ss_handle *h, const void *d_bgr, size_t row_pitch_bytes);
/* Block until all pushed work has completed. */
int ss_sync(ss_handle *h);
/* Block only until every pushed frame has reached the device (host buffers may then be reused). */
int ss_sync_copies(ss_handle *h);

/* Replaces fuse_pyramids' base step up to the histogram (pyramid.py:95-101,67): flushes pending frames,
 * computes each frame's truncated gray base level and its histogram.  counts receives
 * n_frames x num_values int64 entries (num_values = 256 or 65536); may be NULL. */
int ss_finish_begin(ss_handle *h, int64_t *counts);
/* Replaces get_fused_base + collapse + astype (pyramid.py:102-111,57-64,179).  log_prob is the table
 * np.log(probabilities) of PyramidBase.entropy (pyramid.py:68-77), n_frames x num_values float32,
 * computed by the host with numpy so that it is bit-identical to the reference; NULL makes the library
 * compute it with C logf (documented last-ulp differences versus numpy).  out_bgr receives the fused
 * H x W x 3 image in the input sample type. */
int ss_finish_end(ss_handle *h, const float *log_prob, void *out_bgr, size_t out_row_pitch_bytes);
/* Convenience: ss_finish_begin + ss_finish_end(log_prob = NULL). */
int ss_finish(ss_handle *h, void *out_bgr, size_t out_row_pitch_bytes);

/* The integer depth maps (np.argmax results, pyramid.py:51 and :104-105; not returned by the reference
 * API but named by the north-star): level 0..L-1 = per-level Laplacian arg-max; level L = base-level
 * entropy arg-max; level L+1 = base-level deviation arg-max.  int32 frame indices (global, i.e.
 * including frame_index_base), height x width of that level, written to host memory. */
int ss_get_argmax(ss_handle *h, int level, int32_t *out);
/* Fused Laplacian level l (l < L) or fused base (l == L) as H_l x W_l x 3 float32 BGR (tests). */
int ss_get_fused_level(ss_handle *h, int level, float *out);

/* Device-side time of the kernel chain in milliseconds, measured with CUDA events on the library's
 * compute stream: accumulated since the last ss_reset (all batches + finish). */
int ss_kernel_time_ms(ss_handle *h, float *ms);
/* Optional per-launch CUDA-event timing of the fused level kernels: enable, run, then read the summed
 * device time and launch count of one level's kernel (bench.py's roofline figure). */
int ss_set_profiling(ss_handle *h, int on);
int ss_level_kernel_time_ms(ss_handle *h, int level, float *total_ms, int32_t *launches);
/* Number of kernel launches since the last ss_reset. */
int ss_launch_count(ss_handle *h, int64_t *launches);

/* ---- multi-GPU, fused peer exchange (one process per GPU of a box; no collective library, no host in the loop) ----
 * Frames of one stack are sharded over the ranks in contiguous blocks (frame_index_base = first global index).  The
 * rows of every pyramid level are cut into one band per rank.  After its last frame a rank signals READY through an
 * epoch flag in every peer's memory; its merge kernel then reads, for its bands (+ the halo rows its collapse needs),
 * the running values of all ranks over NVLink, takes np.argmax's first maximum (rank order == frame order, strict
 * '>'; pyramid.py:51-54 over the whole stack), pulls the winner's coefficients, collapses its band (pyramid.py:57-64)
 * and stores its rows of the fused image into the root's output buffer (peer stores); the root waits for every
 * band's OUT flag and downloads the image.
 *   setup, once per handle:  ss_peer_export on every rank -> the caller all-gathers the 384-byte blobs (any
 *   transport; shinestacker_b200/rendezvous.py uses a TCP socket) -> ss_peer_connect on every rank.
 *   per stack: push frames as usual, then ss_peer_finish on every rank (out_bgr only meaningful on the root).
 * ss_peer_ptrs / ss_peer_connect_ptrs: the same for several handles inside one process (tests). */
int ss_peer_export(ss_handle *h, void *blob384);
int ss_peer_connect(ss_handle *h, const void *blobs, int world, int rank, int root);
int ss_peer_ptrs(ss_handle *h, void **ptrs6);
int ss_peer_connect_ptrs(ss_handle *h, void *const *ptrs, int world, int rank, int root);
int ss_peer_finish(ss_handle *h, void *out_bgr, size_t out_row_pitch_bytes);
/* The two halves of ss_peer_finish: enqueue this rank's part (returns at once) / wait for it (root: for all bands,
 * then download).  A process that drives several ranks must enqueue all of them before it waits for any. */
int ss_peer_finish_async(ss_handle *h);
int ss_peer_wait(ss_handle *h, void *out_bgr, size_t out_row_pitch_bytes);
/* ss_peer_finish_async in its three phases (1: base selection + READY; 2: wait READY, merge, DONE, banded collapse,
 * OUT; 3: root waits for every OUT).  Only needed when one process drives several ranks on one device (tests): submit
 * phase p of every rank before phase p + 1 of any. */
int ss_peer_phase(ss_handle *h, int phase);
/* Test hook (no GPU needed): the row bands ss_peer_connect derives for an image of this size; out receives
 * 4 x (L + 2) int32 (merge rows [row0, row1) and owned rows [own0, own1) per segment: levels 0..L-1, two base planes),
 * level_h (may be NULL) the L + 1 level heights; returns L. */
int ss_test_peer_bands(int height, int width, int min_size, int world, int rank, int32_t *out, int32_t *level_h);

/* Re-run the kernel chain over the frames still resident in HBM (requires resident_frames >= frames
 * pushed).  Used by bench.py for the "inputs already in HBM" throughput. */
int ss_rerun_resident(ss_handle *h);
/* The two halves of ss_finish_end, separately callable so that a multi-GPU exchange can run between
 * them: base-level entropy/deviation arg-max on this rank's frames; then base blend + collapse +
 * conversion + download. */
int ss_finish_select_base(ss_handle *h, const float *log_prob);
/* Variant for 16-bit stacks: the log-probability of every base pixel's own gray value
 * (n_frames x hL x wL float32) instead of a 65536-entry table per frame; ss_get_base_gray delivers the
 * truncated gray base levels (n_frames x hL x wL int32) the host needs to build it (np.unique per frame,
 * exactly as PyramidBase.entropy does, pyramid.py:67-69). */
int ss_get_base_gray(ss_handle *h, int32_t *out);
/* float_type = 'float-64' (reference base_stack_algo.py:15-16): the same step with the float64 table
 * np.log(probabilities) (n_frames x num_values, or per base pixel when per_px != 0; NULL -> libm log); it
 * also blends the base level.  ss_finish_collapse follows as usual. */
int ss_finish_select_base_f64(ss_handle *h, const double *log_prob, int per_px);
int ss_finish_select_base_px(ss_handle *h, const float *log_prob_px);
int ss_finish_collapse(ss_handle *h, void *out_bgr, size_t out_row_pitch_bytes);
/* The same base step without any host round trip (float-32): PyramidBase.entropy's np.log(probabilities)
 * (pyramid.py:67-69,77) only ever sees the n_base + 1 values float32(c) / n_base, c = 0 .. n_base = hL * wL, so the host
 * supplies that table once (ss_set_log_lut: computed with numpy by the Python binding, libm's logf by default) and
 * the device indexes it with its own histogram.  ss_finish_base_device = ss_finish_begin + base selection, all
 * asynchronous; ss_finish_collapse(h, NULL, 0) then leaves the fused image on the device and ss_download_output
 * fetches it. */
int ss_set_log_lut(ss_handle *h, const float *lut, int n);
int ss_finish_base_device(ss_handle *h);
int ss_download_output(ss_handle *h, void *out_bgr, size_t out_row_pitch_bytes);
/* Time of the i-th timed segment (one per launched batch, then histogram, base select, collapse). */
int ss_segment_time_ms(ss_handle *h, int i, float *ms);

/* ---- test hooks: host arrays out of / through the product kernels, each checked against the
 * reference method it replaces ------------------------------------------------------------------- */
/* Running best energy of level l (after a 1-frame stack: convolve(square(gray(lap))), reference
 * pyramid.py:49-50), H_l x W_l float32. */
int ss_get_best_energy(ss_handle *h, int level, float *out);
/* g_level of one frame as 3 float32 planes (PyramidBase.reduce_layer, pyramid.py:27-32): for
 * level < L the frame index counts within the most recent batch, for level == L it is absolute. */
int ss_get_gauss_level(ss_handle *h, int level, int frame, float *out_planes);
/* One collapse step through the product's collapse kernel (reference pyramid.py:59-63, i.e.
 * expand_layer pyramid.py:34-41 cropped, + layer): planes in, planes out; layer may be NULL (pure
 * expand). */
int ss_test_collapse_step(const float *img_planes, int h2, int w2, const float *layer_planes, int h,
                          int w, double gen_kernel, float *out_planes);

#ifdef __cplusplus
}
#endif
#endif /* SHINESTACK_B200_H */
