// ss_lib.cu -- libshinestack_b200.so: C ABI (include/shinestack_b200.h) + kernel launches.
// sm_100a only; no CPU fallback: every entry point that computes needs a compute-capability-10 GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/shinestack_b200.h"
#include "ss_kernels.cuh"
#include "ss_f64.cuh"

#define SS_EXPORT extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int ss_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define SS_CUDA(call)                                                                           \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return ss_fail(SS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                           __FILE__, __LINE__);                                                 \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
// Tiles that touch an image border do extra work (reflected loads, fix-up passes); they are mapped to the
// lowest block ids so that they start first and the tail of the launch consists of regular tiles.
__device__ __forceinline__ void ss_tile_of_block(int b, int nbx, int nby, int &tx, int &ty)
{
    if (nbx <= 2 || nby <= 2) { tx = b % nbx; ty = b / nbx; return; }
    const int nb = 2 * nbx + 2 * (nby - 2);
    if (b < nbx) { tx = b; ty = 0; }
    else if (b < 2 * nbx) { tx = b - nbx; ty = nby - 1; }
    else if (b < nb) { const int k = b - 2 * nbx; tx = (k & 1) ? nbx - 1 : 0; ty = 1 + (k >> 1); }
    else { const int k = b - nb; tx = 1 + k % (nbx - 2); ty = 1 + k / (nbx - 2); }
}

template <typename InT>
__global__ void __launch_bounds__(SS_NT, 2) ss_level_kernel(const __grid_constant__ SsLevelParams p,
                                                            const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) float ss_smem[];
    int tx, ty;
    ss_tile_of_block(blockIdx.y * gridDim.x + blockIdx.x, gridDim.x, gridDim.y, tx, ty);
    ss_level_tile<InT>(p, &tmap, tx, ty, blockIdx.z,
                       (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x, ss_smem);
}

// Folds the per-group partial states of a frame-split level launch into the running state, groups in
// frame order with strict '>' (== np.argmax's first-maximum rule over the whole stack).
__global__ void ss_merge_groups_kernel(size_t n, int groups, const float *g_e, const int32_t *g_idx,
                                       const float *g_fused, int have_state, float *best_e,
                                       int32_t *best_idx, float *fused)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // pass 1: the winning group (energies only; the loads of a batch of 8 groups are independent)
    float e;
    int win, g0;
    if (have_state) { e = best_e[i]; win = -1; g0 = 0; }
    else { e = g_e[i]; win = 0; g0 = 1; }
#pragma unroll 8
    for (int g = g0; g < groups; ++g) {
        const float eg = g_e[(size_t)g * n + i];
        if (eg > e) { e = eg; win = g; }
    }
    // pass 2: index and coefficients of the winner only
    if (win >= 0) {
        const float *gf = g_fused + 3 * (size_t)win * n;
        best_e[i] = e;
        best_idx[i] = g_idx[(size_t)win * n + i];
        fused[i] = gf[i]; fused[n + i] = gf[n + i]; fused[2 * n + i] = gf[2 * n + i];
    }
}

// img.astype(float32) (reference pyramid.py:126) into planes; only needed when the pyramid has no
// Laplacian level at all (image smaller than 2*min_size) and the input itself is the base level.
template <typename InT>
__global__ void ss_convert_kernel(const InT *in, int pitch_bytes, int h, int w, float *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const InT *s = (const InT *)((const char *)in + (size_t)y * pitch_bytes) + 3 * x;
    const size_t plane = (size_t)h * w, o = (size_t)y * w + x;
    out[o] = (float)s[0];
    out[plane + o] = (float)s[1];
    out[2 * plane + o] = (float)s[2];
}

// Base level, step 1 (reference pyramid.py:99-101 and :67): gray = cvtColor(base).astype(dtype) by
// truncation, plus the per-frame histogram that PyramidBase.entropy obtains from np.unique.
__global__ void ss_base_gray_hist_kernel(const float *base, int n_px, int32_t *gray,
                                         unsigned int *counts, int num_values)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const float *b = base + (size_t)f * 3 * n_px;
    const float v = __fmaf_rn(b[2 * (size_t)n_px + i], 0.299f,
                              __fmaf_rn(b[i], 0.114f, __fmul_rn(b[(size_t)n_px + i], 0.587f)));
    int gv = (int)v;  // C truncation, as ndarray.astype
    gv = gv < 0 ? 0 : (gv >= num_values ? num_values - 1 : gv);
    gray[(size_t)f * n_px + i] = gv;
    atomicAdd(&counts[(size_t)f * num_values + gv], 1u);
}

// numpy's float32 pairwise add.reduce over n contiguous values (reference: ndarray.sum() at pyramid.py:77 and :85;
// SURVEY.md section 9 item 7): up to 128 values eight running sums combined as a fixed tree, above that the run is
// halved (first half rounded down to a multiple of 8) -- DEPTH levels of halving cover n <= 128 << DEPTH.
__device__ __forceinline__ float ss_np_pairwise_leaf(const float *a, int n)
{
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}
template <int DEPTH>
__device__ float ss_np_pairwise_sum(const float *a, int n)
{
    if (n <= 128) return ss_np_pairwise_leaf(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(ss_np_pairwise_sum<DEPTH - 1>(a, n2), ss_np_pairwise_sum<DEPTH - 1>(a + n2, n - n2));
}
template <>
__device__ float ss_np_pairwise_sum<0>(const float *a, int n) { return ss_np_pairwise_leaf(a, n); }

#define SS_MAX_WINDOW 31         /* base-level window side: kernel_size up to 32 (the oracle's range) */
#define SS_SMALL_WINDOW 11       /* sides up to this use the instantiation with the small per-thread window buffers */

// Base level, step 2a (reference pyramid.py:66-93): per (pixel, frame) the window entropy
// -sum(v*log p[v]) and the window variance, both in numpy's evaluation order.
// log p[v] comes in one of three forms: per_px == 0: a per-frame table indexed by the value; per_px == 1: the
// log-probability of every base pixel's own value, indexed by the pixel; per_px == 2: lut[counts[f][v]], lut being
// the table log(float32(c) / n_px) for every possible count c = 0 .. n_px (computed once by the host -- with numpy
// when Python drives the library, so that it is numpy's log -- which removes the host round trip of the histogram).
template <int MAXSIDE>
__global__ void ss_base_metrics_kernel(const int32_t *gray, const float *log_p, int per_px, int h, int w,
                                       int num_values, int side, float *ent_out, float *dev_out,
                                       const unsigned int *counts = nullptr)
{
    constexpr int DEPTH = MAXSIDE * MAXSIDE <= 128 ? 0 : 3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    const int n_px = h * w;
    if (i >= n_px) return;
    const int y = i / w, x = i - y * w;
    const int pad = (side - 1) / 2, n = side * side;
    float buf[MAXSIDE * MAXSIDE];
    int vals[MAXSIDE * MAXSIDE];
    const int32_t *g = gray + (size_t)f * n_px;
    // log p[v]: either a per-frame table indexed by the value, or (16-bit stacks) the log-probability of
    // every base pixel's own value, indexed by the pixel
    const float *lp = per_px == 2 ? log_p : log_p + (size_t)f * (per_px ? n_px : num_values);
    const unsigned int *cnt = counts + (size_t)f * num_values;
    int c = 0;
    long long s = 0;
    for (int dy = -pad; dy <= pad; ++dy) {
        int yy = y + dy;
        while (yy < 0 || yy >= h) yy = yy < 0 ? -yy : 2 * h - 2 - yy;
        for (int dx = -pad; dx <= pad; ++dx) {
            int xx = x + dx;
            while (xx < 0 || xx >= w) xx = xx < 0 ? -xx : 2 * w - 2 - xx;
            const int v = g[yy * w + xx];
            vals[c] = v;
            buf[c] = __fmul_rn((float)v, per_px == 2 ? lp[cnt[v]] : lp[per_px ? yy * w + xx : v]);
            s += v;
            ++c;
        }
    }
    ent_out[(size_t)f * n_px + i] = __fmul_rn(-1.f, ss_np_pairwise_sum<DEPTH>(buf, n));
    const float mean = (float)((double)s / (double)n);
    for (int k = 0; k < n; ++k) {
        const float d = __fsub_rn((float)vals[k], mean);
        buf[k] = __fmul_rn(d, d);
    }
    dev_out[(size_t)f * n_px + i] = __fdiv_rn(ss_np_pairwise_sum<DEPTH>(buf, n), (float)n);
}

// Base level, step 2b (reference pyramid.py:102-105): running strict-'>' arg-max of entropy and of
// deviation over the frames, and the base pixels of the two winners.
__global__ void ss_base_select_kernel(const float *ent, const float *dev, const float *base, int n_px,
                                      int n_frames, int frame_index0, float *best_ve, int32_t *best_ie,
                                      float *best_vd, int32_t *best_id, float *pix_e, float *pix_d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    float ve = ent[i], vd = dev[i];
    int ie = 0, id = 0;
    for (int f = 1; f < n_frames; ++f) {
        const float e = ent[(size_t)f * n_px + i], d = dev[(size_t)f * n_px + i];
        if (e > ve) { ve = e; ie = f; }
        if (d > vd) { vd = d; id = f; }
    }
    best_ve[i] = ve; best_ie[i] = frame_index0 + ie;
    best_vd[i] = vd; best_id[i] = frame_index0 + id;
    for (int ch = 0; ch < 3; ++ch) {
        pix_e[(size_t)ch * n_px + i] = base[((size_t)ie * 3 + ch) * n_px + i];
        pix_d[(size_t)ch * n_px + i] = base[((size_t)id * 3 + ch) * n_px + i];
    }
}

// get_fused_base's blend (reference pyramid.py:106-111): ((0 + base[best_e]) + base[best_d]) / 2.
__global__ void ss_base_fuse_kernel(const float *pix_e, const float *pix_d, size_t n, float *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = __fdiv_rn(__fadd_rn(__fadd_rn(0.f, pix_e[i]), pix_d[i]), 2.f);
}

struct SsCollapseParams {
    const float *img;    // level l+1 image, 3 planes h2 x w2
    const float *layer;  // fused Laplacian of level l, 3 planes h x w (may be null: expand only)
    float *out;          // 3 planes h x w (when !final)
    void *out_px;        // final: interleaved uint8/uint16 H x W x 3
    int out_pitch;       // bytes
    int h, w, h2, w2;
    int final_bytes;     // 0: not final; 1 / 2: clip(abs()) and truncate to uint8 / uint16
    int qy0;             // first quad row of the launch (multi-GPU: every rank collapses a band of rows)
    int y_end;           // rows >= y_end are not produced (0: all)
    float k25[25];
};

// One collapse step (reference pyramid.py:59-64): img = expand(img)[:h,:w] + layer, the expand being
// the zero-skipping form of expand_layer (pyramid.py:34-41); 4*acc + layer is one fma (4*acc is exact);
// the last step also applies clip(abs(img), 0, max) and the truncating cast of pyramid.py:179.
// One thread per aligned 2x2 output quad: its taps are a 3x3 window of the coarser image.
__device__ __forceinline__ float ss_collapse_px_generic(const float *g, int h2, int w2, int y, int x,
                                                        const float *k25, bool tail)
{
    const int wz = 2 * w2, hz = 2 * h2;
    float acc = 0.f;
    for (int dy = -2; dy <= 2; ++dy) {
        int zy = y + dy;
        if (zy < 0) zy = -zy;
        if (zy >= hz) zy = 2 * hz - 2 - zy;
        if (zy & 1) continue;
        const float *row = g + (size_t)(zy >> 1) * w2;
        for (int dx = -2; dx <= 2; ++dx) {
            int zx = x + dx;
            if (zx < 0) zx = -zx;
            if (zx >= wz) zx = 2 * wz - 2 - zx;
            if (zx & 1) continue;
            const float t = row[zx >> 1], k = k25[(dy + 2) * 5 + dx + 2];
            acc = tail ? __fadd_rn(acc, __fmul_rn(t, k)) : __fmaf_rn(t, k, acc);
        }
    }
    return acc;
}

__global__ void __launch_bounds__(256) ss_collapse_kernel(const __grid_constant__ SsCollapseParams p)
{
    const int qx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int qy = p.qy0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    const int x = 2 * qx, y = 2 * qy;
    if (x >= p.w || y >= p.h || (p.y_end > 0 && y >= p.y_end)) return;
    const int wz = 2 * p.w2, wz_vec = wz - (wz % 8);
    const bool tail = x + 1 >= wz_vec;
    float v[3][4];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float *g = p.img + (size_t)ch * p.h2 * p.w2;
        float acc[4];
        if (!tail) {
            // window rows (y-2)/2 + a, columns (x-2)/2 + b; out-of-image indices as REFLECT101 on the
            // zero-stuffed image sees them: -1 -> 1, h2 -> h2-1
            float win[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                int I = qy - 1 + a;
                I = I < 0 ? -I : (I >= p.h2 ? 2 * p.h2 - 1 - I : I);
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    int J = qx - 1 + b;
                    J = J < 0 ? -J : (J >= p.w2 ? 2 * p.w2 - 1 - J : J);
                    win[a][b] = g[(size_t)I * p.w2 + J];
                }
            }
            ss_expand_block<2>(win, p.k25, acc);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                acc[q] = ss_collapse_px_generic(g, p.h2, p.w2, y + (q >> 1), x + (q & 1), p.k25, x + (q & 1) >= wz_vec);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int yy = y + (q >> 1), xx = x + (q & 1);
            float e;
            if (p.layer && yy < p.h && xx < p.w) e = __fmaf_rn(4.f, acc[q], p.layer[((size_t)ch * p.h + yy) * p.w + xx]);
            else e = __fmul_rn(4.f, acc[q]);
            v[ch][q] = e;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int yy = y + (q >> 1), xx = x + (q & 1);
        if (yy >= p.h || xx >= p.w || (p.y_end > 0 && yy >= p.y_end)) continue;
        if (!p.final_bytes) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) p.out[((size_t)ch * p.h + yy) * p.w + xx] = v[ch][q];
        } else {
            const float maxv = p.final_bytes == 1 ? 255.f : 65535.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float a = fabsf(v[ch][q]);
                a = a > maxv ? maxv : a;   // abs() is never negative; NaN cannot occur
                const unsigned int u = __float2uint_rz(a);
                if (p.final_bytes == 1) ((uint8_t *)((char *)p.out_px + (size_t)yy * p.out_pitch))[3 * xx + ch] = (uint8_t)u;
                else ((uint16_t *)((char *)p.out_px + (size_t)yy * p.out_pitch))[3 * xx + ch] = (uint16_t)u;
            }
        }
    }
}

// clip(abs()) + cast of a planar image without any expand (pyramid with no Laplacian level).
__global__ void ss_finalize_kernel(const float *img, int h, int w, int bytes, void *out, int pitch, int y_first = 0)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = y_first + blockIdx.y;
    if (x >= w) return;
    const float maxv = bytes == 1 ? 255.f : 65535.f;
    for (int ch = 0; ch < 3; ++ch) {
        float a = fabsf(img[((size_t)ch * h + y) * w + x]);
        a = a > maxv ? maxv : a;
        const unsigned int u = __float2uint_rz(a);
        if (bytes == 1) ((uint8_t *)((char *)out + (size_t)y * pitch))[3 * x + ch] = (uint8_t)u;
        else ((uint16_t *)((char *)out + (size_t)y * pitch))[3 * x + ch] = (uint16_t)u;
    }
}

// The running state of all levels and of the base level lives in three contiguous arrays (values, indices, 3-plane
// coefficients) in one fixed order, so that the multi-GPU exchange addresses every level with one set of offsets.
#define SS_MAX_SEGMENTS 16   // pyramid levels + 2 base planes: up to 14 Laplacian levels (a 64k x 64k image at min_size 4)
struct SsSegments {        // per plane-group (pyramid level / base plane): pixel offset, pixel count, coefficient offset
    int count;
    unsigned long long px_off[SS_MAX_SEGMENTS], n_px[SS_MAX_SEGMENTS], coef_off[SS_MAX_SEGMENTS];
};

// ---- fused peer exchange over NVLink (one process per GPU, no collective library) ---------------------
// Every rank has built the running state of ITS frames (contiguous block of the stack) for the whole image.  The
// image rows of every pyramid level are cut into one band per rank; a rank merges its bands (plus the few halo rows
// its collapse needs) straight out of the peers' memory -- arg-max over the ranks in frame order with the strict '>'
// of np.argmax, then the winner's three coefficients -- collapses its band of the pyramid and stores its rows of the
// fused image into the root's output buffer.  Synchronisation is by epoch flags in peer memory.
#define SS_MAX_RANKS 16
struct SsPeerState {
    int world, rank;
    const float *val[SS_MAX_RANKS];       // running values (energies / base entropy / deviation) of every rank
    const int32_t *idx[SS_MAX_RANKS];     // global frame indices
    const float *coef[SS_MAX_RANKS];      // winners' coefficients (3 planes per segment)
};
struct SsBandRows {                       // per segment: rows this rank merges, and the rows it owns (depth maps)
    int count;
    int row0[SS_MAX_SEGMENTS], row1[SS_MAX_SEGMENTS], own0[SS_MAX_SEGMENTS], own1[SS_MAX_SEGMENTS], width[SS_MAX_SEGMENTS];
};
enum { SS_FLAG_READY = 0, SS_FLAG_DONE = 1, SS_FLAG_OUT = 2 };

__global__ void ss_flag_signal_kernel(unsigned int *const *peer_flags, int world, int kind, int self, unsigned int epoch,
                                      int only)
{
    const int p = threadIdx.x;
    if (p >= world || (only >= 0 && p != only)) return;
    __threadfence_system();   // everything this device wrote before (earlier kernels of the stream) is visible first
    volatile unsigned int *f = peer_flags[p] + kind * SS_MAX_RANKS + self;
    *f = epoch;
    __threadfence_system();
}

__device__ __forceinline__ void ss_flag_wait(const unsigned int *flags, int world, int kind, unsigned int epoch)
{
    if (threadIdx.x < world) {
        const volatile unsigned int *f = flags + kind * SS_MAX_RANKS + threadIdx.x;
        while ((int)(*f - epoch) < 0) __nanosleep(64);
    }
    __threadfence_system();
    __syncthreads();
}

__global__ void ss_flag_wait_kernel(const unsigned int *flags, int world, int kind, unsigned int epoch)
{
    ss_flag_wait(flags, world, kind, epoch);
}

// One thread per pixel of the rows [row0, row1) of every segment.
__global__ void __launch_bounds__(256) ss_peer_merge_kernel(const __grid_constant__ SsPeerState ps,
                                                            const __grid_constant__ SsSegments seg,
                                                            const __grid_constant__ SsBandRows band,
                                                            const unsigned int *flags, unsigned int epoch,
                                                            float *mrg_coef, int32_t *root_idx)
{
    (void)flags; (void)epoch;   // (the READY wait is a one-CTA kernel launched just before this one: a grid of
                                // spinning CTAs could starve the very kernels it waits for when ranks share a GPU)
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int s = 0;
    for (; s < band.count; ++s) {
        const size_t n = (size_t)(band.row1[s] - band.row0[s]) * band.width[s];
        if (i < n) break;
        i -= n;
    }
    if (s >= band.count) return;
    const int row = band.row0[s] + (int)(i / band.width[s]);
    const size_t px = (size_t)row * band.width[s] + (i % band.width[s]);   // pixel inside the segment
    const size_t k = seg.px_off[s] + px;
    float best = ps.val[0][k];
    int win = 0;
    for (int r = 1; r < ps.world; ++r) {          // rank order == frame order: strict '>' keeps the first maximum
        const float v = ps.val[r][k];
        if (v > best) { best = v; win = r; }
    }
    const size_t c = seg.coef_off[s] + px, np_ = seg.n_px[s];
    const float *src = ps.coef[win];
    mrg_coef[c] = src[c]; mrg_coef[c + np_] = src[c + np_]; mrg_coef[c + 2 * np_] = src[c + 2 * np_];
    if (root_idx && row >= band.own0[s] && row < band.own1[s]) root_idx[k] = ps.idx[win][k];
}

// ---------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------
struct SsLevel {
    int h, w;
    float *best_e = nullptr;
    int32_t *best_idx = nullptr;
    float *fused = nullptr;  // 3 planes
    float *gnext = nullptr;  // scratch: g_{l+1} of the frames of one batch (unused for the last level)
    int gpitch = 0;          // row pitch of gnext in floats (multiple of 4: 16-byte aligned rows)
    int max_groups = 1;      // frame-split factor for a full batch (levels with too few tiles to fill the GPU)
    float *grp_e = nullptr;  // per-group partial states (max_groups slices) when max_groups > 1
    int32_t *grp_idx = nullptr;
    float *grp_fused = nullptr;
};

struct ss_handle {
    ss_config cfg;
    int L = 0;                       // number of Laplacian levels
    std::vector<int> lh, lw;         // sizes of levels 0..L
    std::vector<SsLevel> lv;         // 0..L-1
    float k25[25];
    int side = 5;                    // base window side
    int num_values = 256;
    // frame store (ring of n_slots device frames)
    char *frames = nullptr;
    size_t frame_pitch = 0, frame_bytes = 0, frame_shift = 0;
    float *lap_scratch = nullptr;    // float32 levels: 3 x 64 x 64 floats per CTA of the largest level >= 1 launch
    size_t lap_scratch_ctas = 0;
    CUtensorMap tmap_frames;         // uint8 frame store as a 2-D byte tensor [n_slots * H rows][pitch]: TMA boxes
    int n_slots = 0, batch = 0;
    // base level
    float *base = nullptr;           // max_frames x 3 x hL x wL
    int32_t *gray = nullptr;         // max_frames x hL x wL
    unsigned int *counts = nullptr;  // max_frames x num_values
    float *log_lut = nullptr;        // n_base + 1 entries: log(float32(c) / n_base), see ss_set_log_lut
    float *log_p = nullptr;          // max_frames x num_values
    // contiguous running state in exchange order: [level 0 .. L-1 | base entropy | base deviation]
    float *state_val = nullptr;      // n_keys floats
    int32_t *state_idx = nullptr;    // n_keys
    float *state_coef = nullptr;     // n_coeffs floats: 3 planes per segment
    SsSegments seg;
    float *best_ve = nullptr, *best_vd = nullptr, *pix_e = nullptr, *pix_d = nullptr;   // views into the above
    float *ent = nullptr, *dev = nullptr;  // max_frames x hL x wL window entropy / variance
    int32_t *best_ie = nullptr, *best_id = nullptr;
    float *fused_base = nullptr;     // 3 x hL x wL
    float *tmp_a = nullptr, *tmp_b = nullptr;  // collapse ping-pong, 3 x H x W / 4 each... see alloc
    void *out_dev = nullptr;         // interleaved output image
    size_t out_pitch = 0;
    // fused peer exchange (ss_peer_*)
    bool peer_on = false;
    int peer_world = 1, peer_rank = 0, peer_root = 0;
    unsigned int peer_epoch = 0;
    unsigned int *flags = nullptr;               // [3][SS_MAX_RANKS] epoch flags, written by the peers
    unsigned int **d_peer_flags = nullptr;       // device array of every rank's flags pointer
    void *peer_open[SS_MAX_RANKS][6] = {};       // IPC-opened allocations of the peers (closed in ss_destroy)
    SsPeerState pstate;
    SsBandRows band;
    float *mrg_coef = nullptr;                   // merged coefficients (exchange layout), rows of `band` only
    int32_t *mrg_idx = nullptr;                  // root: merged arg-max indices of the whole stack
    int32_t *root_mrg_idx = nullptr;             // that buffer as seen from this rank
    void *root_out = nullptr;                    // the root's output image as seen from this rank
    size_t n_keys = 0, n_coeffs = 0;
    // streams / events
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    std::vector<cudaEvent_t> ev_copied, ev_done;  // per batch slot
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timers;
    size_t timers_used = 0;
    bool profile = false;            // per-launch events around every level kernel
    struct LevelTimer { int level; cudaEvent_t a, b; };
    std::vector<LevelTimer> ltimers;
    size_t ltimers_used = 0;
    // progress
    int n_pushed = 0;      // frames handed over
    int n_launched = 0;    // frames whose level chain has been launched
    bool finished_begin = false;
    bool have_state = false;
    long long launches = 0;
    // ---- float_type = 'float-64' path (ss_f64.cuh): one frame at a time, all state in global memory -----
    bool f64 = false;
    Ss64Kernel k64;
    std::vector<double *> g64;        // Gaussian pyramid of the current frame, levels 0..L-1 (3 planes each)
    std::vector<double *> fused64;    // running winner coefficients per level (3 planes)
    double *lap64 = nullptr;          // Laplacian of the current frame / level (3 planes of level 0)
    float *sq64 = nullptr;            // gray^2 plane
    double *base64 = nullptr;         // max_frames x 3 x hL x wL
    double *logp64 = nullptr, *ent64 = nullptr, *dev64 = nullptr, *fusedbase64 = nullptr;
    double *tmp64_a = nullptr, *tmp64_b = nullptr;
};

static int ss_levels_for(int h, int w, int min_size, std::vector<int> &lh, std::vector<int> &lw)
{
    // reference pyramid.py:165 and :127-131
    lh.assign(1, h);
    lw.assign(1, w);
    const double ratio = (double)(h < w ? h : w) / (double)min_size;
    int levels = ratio > 0 ? (int)log2(ratio) : 0;  // int() truncates toward zero
    if (levels < 0) levels = 0;
    for (int i = 0; i < levels; ++i) {
        const int nh = (lh.back() + 1) / 2, nw = (lw.back() + 1) / 2;
        if ((nh < nw ? nh : nw) < 4) break;
        lh.push_back(nh);
        lw.push_back(nw);
    }
    return (int)lh.size() - 1;
}

static void ss_make_kernel(double a, float *k25)
{
    // reference pyramid.py:19-22 (float64 outer product, converted to float32 by cv2.filter2D)
    const double k[5] = {0.25 - a / 2.0, 0.25, a, 0.25, 0.25 - a / 2.0};
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) k25[i * 5 + j] = (float)(k[i] * k[j]);
}

static int ss_timer_begin(ss_handle *h)
{
    if (h->timers_used == h->timers.size()) {
        cudaEvent_t a, b;
        SS_CUDA(cudaEventCreate(&a));
        SS_CUDA(cudaEventCreate(&b));
        h->timers.push_back({a, b});
    }
    SS_CUDA(cudaEventRecord(h->timers[h->timers_used].first, h->s_comp));
    return 0;
}
static int ss_timer_end(ss_handle *h)
{
    SS_CUDA(cudaEventRecord(h->timers[h->timers_used].second, h->s_comp));
    h->timers_used++;
    return 0;
}

SS_EXPORT const char *ss_version(void) { return "shinestack_b200 0.1 (sm_100a)"; }
SS_EXPORT const char *ss_last_error(void) { return g_err; }

SS_EXPORT int ss_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess &&
            major == 10)
            ++ok;
    }
    return ok;
}

template <typename T>
static int ss_dalloc(T **p, size_t count)
{
    SS_CUDA(cudaMalloc((void **)p, count * sizeof(T) + 16));
    return 0;
}

SS_EXPORT void ss_destroy(ss_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->s_comp) cudaStreamSynchronize(h->s_comp);
    if (h->s_copy) cudaStreamSynchronize(h->s_copy);
    for (auto &l : h->lv) {
        cudaFree(l.gnext);
        cudaFree(l.grp_e); cudaFree(l.grp_idx); cudaFree(l.grp_fused);
    }
    cudaFree(h->lap_scratch);
    cudaFree(h->frames); cudaFree(h->base); cudaFree(h->gray); cudaFree(h->counts);
    cudaFree(h->log_lut);
    cudaFree(h->log_p); cudaFree(h->state_val); cudaFree(h->state_idx); cudaFree(h->state_coef);
    cudaFree(h->ent); cudaFree(h->dev);
    cudaFree(h->fused_base);
    for (auto &row : h->peer_open)
        for (void *pm : row)
            if (pm) cudaIpcCloseMemHandle(pm);
    cudaFree(h->flags); cudaFree(h->d_peer_flags); cudaFree(h->mrg_coef); cudaFree(h->mrg_idx);
    cudaFree(h->tmp_a); cudaFree(h->tmp_b); cudaFree(h->out_dev);
    for (auto e : h->ev_copied) cudaEventDestroy(e);
    for (auto e : h->ev_done) cudaEventDestroy(e);
    for (auto &t : h->timers) { cudaEventDestroy(t.first); cudaEventDestroy(t.second); }
    for (auto &t : h->ltimers) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto p64 : h->g64) cudaFree(p64);
    for (auto p64 : h->fused64) cudaFree(p64);
    cudaFree(h->lap64); cudaFree(h->sq64); cudaFree(h->base64); cudaFree(h->logp64); cudaFree(h->ent64);
    cudaFree(h->dev64); cudaFree(h->fusedbase64); cudaFree(h->tmp64_a); cudaFree(h->tmp64_b);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    if (h->s_comp) cudaStreamDestroy(h->s_comp);
    delete h;
}

static int ss_create_impl(const ss_config *cfg, ss_handle *h)
{
    h->cfg = *cfg;
    SS_CUDA(cudaSetDevice(cfg->device));
    int major = 0;
    SS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cfg->device));
    if (major != 10)
        return ss_fail(SS_ERR_NO_DEVICE, "device %d has compute capability %d.x; this library is "
                       "built for sm_100a only and has no fallback path", cfg->device, major);
    h->L = ss_levels_for(cfg->height, cfg->width, cfg->min_size, h->lh, h->lw);
    ss_make_kernel(cfg->gen_kernel, h->k25);
    h->side = 2 * ((cfg->kernel_size - 1) / 2) + 1;  // len(arange(-pad, pad+1)), pyramid.py:18,120
    h->num_values = cfg->bytes_per_sample == 1 ? 256 : 65536;
    h->batch = cfg->float_bits == 64 ? 1 : (cfg->batch_frames > 0 ? cfg->batch_frames : 8);   // float-64: frame by frame
    if (h->batch > cfg->max_frames) h->batch = cfg->max_frames;
    // frame ring: all frames resident when they fit the request, else two batches
    h->n_slots = cfg->resident_frames > 0 ? cfg->resident_frames : 2 * h->batch;
    if (h->n_slots < h->batch) h->n_slots = h->batch;
    h->n_slots = (h->n_slots + h->batch - 1) / h->batch * h->batch;
    // uint8: the whole store starts 8 bytes into its allocation and every row pitch is a multiple of 16, so every
    // row starts at 8 mod 16: the byte column of pixel x0-8 (x0 a multiple of 64) in the tensor map is then a
    // multiple of 16, which TMA requires of a box origin.  Rows stay contiguous when W*3 is already a multiple of
    // 16 (one 1-D DMA per frame).
    h->frame_shift = cfg->bytes_per_sample == 1 ? 8 : 0;
    h->frame_pitch = ((size_t)cfg->width * 3 * cfg->bytes_per_sample + 15) / 16 * 16;
    h->frame_bytes = h->frame_pitch * cfg->height;
    SS_CUDA(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    SS_CUDA(cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking));
    const int n_bslots = h->n_slots / h->batch;
    h->ev_copied.resize(n_bslots);
    h->ev_done.resize(n_bslots);
    for (int i = 0; i < n_bslots; ++i) {
        SS_CUDA(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
        SS_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
    int rc;
    if ((rc = ss_dalloc(&h->frames, h->frame_bytes * h->n_slots + 16))) return rc;
    memset(&h->tmap_frames, 0, sizeof(h->tmap_frames));
    if (cfg->bytes_per_sample == 1) {
        typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
            return ss_fail(SS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        const cuuint64_t gdim[2] = {(cuuint64_t)h->frame_pitch, (cuuint64_t)cfg->height * h->n_slots};
        const cuuint64_t gstride[1] = {(cuuint64_t)h->frame_pitch};
        const cuuint32_t box[2] = {SS_RAWS, SS_P0H};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = ((encode_fn)fn)(&h->tmap_frames, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, h->frames, gdim, gstride,
                                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return ss_fail(SS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    }
    if (h->L + 2 > SS_MAX_SEGMENTS)
        return ss_fail(SS_ERR_SHAPE, "pyramid with %d levels is not supported (max %d)", h->L, SS_MAX_SEGMENTS - 2);
    h->lv.resize(h->L);
    {
        // segments of the contiguous state: every level, then the two base planes
        const size_t nbase = (size_t)h->lh[h->L] * h->lw[h->L];
        memset(&h->seg, 0, sizeof(h->seg));
        size_t px = 0, co = 0;
        for (int l = 0; l < h->L + 2; ++l) {
            const size_t n = l < h->L ? (size_t)h->lh[l] * h->lw[l] : nbase;
            h->seg.px_off[l] = px; h->seg.n_px[l] = n; h->seg.coef_off[l] = co;
            px += n; co += 3 * n;
        }
        h->seg.count = h->L + 2;
        h->n_keys = px; h->n_coeffs = co;
        if ((rc = ss_dalloc(&h->state_val, px))) return rc;
        if ((rc = ss_dalloc(&h->state_idx, px))) return rc;
        if ((rc = ss_dalloc(&h->state_coef, co))) return rc;
        h->best_ve = h->state_val + h->seg.px_off[h->L];     h->best_ie = h->state_idx + h->seg.px_off[h->L];
        h->best_vd = h->state_val + h->seg.px_off[h->L + 1]; h->best_id = h->state_idx + h->seg.px_off[h->L + 1];
        h->pix_e = h->state_coef + h->seg.coef_off[h->L];    h->pix_d = h->state_coef + h->seg.coef_off[h->L + 1];
    }
    for (int l = 0; l < h->L; ++l) {
        SsLevel &v = h->lv[l];
        v.h = h->lh[l]; v.w = h->lw[l];
        const size_t n = (size_t)v.h * v.w;
        v.best_e = h->state_val + h->seg.px_off[l];
        v.best_idx = h->state_idx + h->seg.px_off[l];
        v.fused = h->state_coef + h->seg.coef_off[l];
        v.gpitch = (h->lw[l + 1] + 3) / 4 * 4;
        {
            // frame-split so that every level launch has >= ~6 waves of CTAs (2 CTAs/SM x 148 SMs per wave): fewer
            // waves leave SMs idle at the tail, more groups cost extra first-frame stores and merge work
            const int tiles = ((v.w + SS_TX - 1) / SS_TX) * ((v.h + SS_TY - 1) / SS_TY);
            const char *wenv = getenv("SS_TARGET_WAVES");   // tuning knob (default 6)
            const int waves = wenv && atoi(wenv) > 0 ? atoi(wenv) : 6;   // measured best on B200 (sweep 6..28)
            int g = (waves * 2 * 148 + tiles - 1) / tiles;
            if (g < 1) g = 1;
            if (g > h->batch) g = h->batch;
            v.max_groups = g;
            if (g > 1) {
                if ((rc = ss_dalloc(&v.grp_e, (size_t)g * n))) return rc;
                if ((rc = ss_dalloc(&v.grp_idx, (size_t)g * n))) return rc;
                if ((rc = ss_dalloc(&v.grp_fused, (size_t)g * 3 * n))) return rc;
            }
        }
        if (l + 1 < h->L)
            if ((rc = ss_dalloc(&v.gnext, (size_t)h->batch * 3 * h->lh[l + 1] * v.gpitch))) return rc;
    }
    for (int l = 1; l < h->L; ++l) {
        const size_t ctas = (size_t)((h->lv[l].w + SS_TX - 1) / SS_TX) * ((h->lv[l].h + SS_TY - 1) / SS_TY) * h->lv[l].max_groups;
        if (ctas > h->lap_scratch_ctas) h->lap_scratch_ctas = ctas;
    }
    if (h->lap_scratch_ctas && (rc = ss_dalloc(&h->lap_scratch, h->lap_scratch_ctas * 3 * SS_TX * SS_TY))) return rc;
    const size_t nb = (size_t)h->lh[h->L] * h->lw[h->L];
    if ((rc = ss_dalloc(&h->base, (size_t)cfg->max_frames * 3 * nb))) return rc;
    if ((rc = ss_dalloc(&h->gray, (size_t)cfg->max_frames * nb))) return rc;
    if ((rc = ss_dalloc(&h->counts, (size_t)cfg->max_frames * h->num_values))) return rc;
    if ((rc = ss_dalloc(&h->log_p, (size_t)cfg->max_frames * ((size_t)h->num_values > nb ? (size_t)h->num_values : nb)))) return rc;
    {
        // default count -> log-probability table from libm; ss_set_log_lut replaces it (numpy's log from Python)
        if ((rc = ss_dalloc(&h->log_lut, nb + 1))) return rc;
        std::vector<float> lut(nb + 1);
        for (size_t c = 0; c <= nb; ++c)
            lut[c] = c ? logf((float)((double)(float)c / (double)nb)) : -INFINITY;
        SS_CUDA(cudaMemcpy(h->log_lut, lut.data(), (nb + 1) * sizeof(float), cudaMemcpyHostToDevice));
    }
    if ((rc = ss_dalloc(&h->ent, (size_t)cfg->max_frames * nb))) return rc;
    if ((rc = ss_dalloc(&h->dev, (size_t)cfg->max_frames * nb))) return rc;
    if ((rc = ss_dalloc(&h->fused_base, 3 * nb))) return rc;
    if (h->L >= 2) {
        // collapse intermediates: the largest one is level 1
        const size_t n1 = 3 * (size_t)h->lh[1] * h->lw[1];
        if ((rc = ss_dalloc(&h->tmp_a, n1))) return rc;
        if ((rc = ss_dalloc(&h->tmp_b, n1))) return rc;
    }
    h->out_pitch = (size_t)cfg->width * 3 * cfg->bytes_per_sample;
    if ((rc = ss_dalloc((char **)&h->out_dev, h->out_pitch * cfg->height))) return rc;
    if (cfg->float_bits == 64) {
        h->f64 = true;
        const double a = cfg->gen_kernel;
        const double k1[5] = {0.25 - a / 2.0, 0.25, a, 0.25, 0.25 - a / 2.0};
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                h->k64.k[i * 5 + j] = k1[i] * k1[j];
                h->k64.kf[i * 5 + j] = (float)(k1[i] * k1[j]);
            }
        h->g64.assign(h->L + 1, nullptr);
        h->fused64.assign(h->L, nullptr);
        for (int l = 0; l <= h->L; ++l)
            if ((rc = ss_dalloc(&h->g64[l], 3 * (size_t)h->lh[l] * h->lw[l]))) return rc;
        for (int l = 0; l < h->L; ++l)
            if ((rc = ss_dalloc(&h->fused64[l], 3 * (size_t)h->lh[l] * h->lw[l]))) return rc;
        const size_t n0 = (size_t)h->lh[0] * h->lw[0];
        if ((rc = ss_dalloc(&h->lap64, 3 * n0))) return rc;
        if ((rc = ss_dalloc(&h->sq64, n0))) return rc;
        if ((rc = ss_dalloc(&h->base64, (size_t)cfg->max_frames * 3 * nb))) return rc;
        if ((rc = ss_dalloc(&h->logp64, (size_t)cfg->max_frames * ((size_t)h->num_values > nb ? (size_t)h->num_values : nb)))) return rc;
        if ((rc = ss_dalloc(&h->ent64, (size_t)cfg->max_frames * nb))) return rc;
        if ((rc = ss_dalloc(&h->dev64, (size_t)cfg->max_frames * nb))) return rc;
        if ((rc = ss_dalloc(&h->fusedbase64, 3 * nb))) return rc;
        if (h->L >= 2) {
            const size_t n1 = 3 * (size_t)h->lh[1] * h->lw[1];
            if ((rc = ss_dalloc(&h->tmp64_a, n1))) return rc;
            if ((rc = ss_dalloc(&h->tmp64_b, n1))) return rc;
        }
    }
    SS_CUDA(cudaFuncSetAttribute(ss_level_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)SS_SMEM_BYTES));
    SS_CUDA(cudaFuncSetAttribute(ss_level_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)SS_SMEM_BYTES));
    SS_CUDA(cudaFuncSetAttribute(ss_level_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)SS_SMEM_BYTES));
    return 0;
}

SS_EXPORT int ss_create(const ss_config *cfg, ss_handle **out)
{
    if (!cfg || !out) return ss_fail(SS_ERR_INVALID_OPTION, "null argument");
    *out = nullptr;
    if (cfg->float_bits != 32 && cfg->float_bits != 64)
        return ss_fail(SS_ERR_INVALID_OPTION, "float_type: valid values are float-32 and float-64 (got %d bits)",
                       cfg->float_bits);
    if (cfg->bytes_per_sample != 1 && cfg->bytes_per_sample != 2)
        return ss_fail(SS_ERR_BIT_DEPTH, "bytes_per_sample must be 1 or 2 (got %d)", cfg->bytes_per_sample);
    if (cfg->width < 4 || cfg->height < 4)
        return ss_fail(SS_ERR_SHAPE, "image %dx%d too small", cfg->width, cfg->height);
    if (cfg->min_size < 1) return ss_fail(SS_ERR_INVALID_OPTION, "min_size must be >= 1");
    if (cfg->kernel_size < 1 || 2 * ((cfg->kernel_size - 1) / 2) + 1 > SS_MAX_WINDOW)
        return ss_fail(SS_ERR_INVALID_OPTION, "kernel_size must be in 1..%d", SS_MAX_WINDOW + 1);
    if (cfg->max_frames < 1) return ss_fail(SS_ERR_INVALID_OPTION, "max_frames must be >= 1");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return ss_fail(SS_ERR_NO_DEVICE, "no CUDA device visible; libshinestack_b200 has no CPU "
                       "fallback");
    }
    if (cfg->device < 0 || cfg->device >= n)
        return ss_fail(SS_ERR_NO_DEVICE, "device %d out of range (%d visible)", cfg->device, n);
    ss_handle *h = new (std::nothrow) ss_handle();
    if (!h) return ss_fail(SS_ERR_NOMEM, "out of host memory");
    const int rc = ss_create_impl(cfg, h);
    if (rc) {
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        ss_destroy(h);
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    *out = h;
    return 0;
}

SS_EXPORT int ss_reset(ss_handle *h)
{
    if (!h) return ss_fail(SS_ERR_STATE, "null handle");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    SS_CUDA(cudaStreamSynchronize(h->s_copy));
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    h->n_pushed = h->n_launched = 0;
    h->finished_begin = false;
    h->have_state = false;
    h->timers_used = 0;
    h->ltimers_used = 0;
    h->launches = 0;
    return 0;
}

SS_EXPORT int ss_num_levels(const ss_handle *h) { return h ? h->L : SS_ERR_STATE; }

SS_EXPORT int ss_level_size(const ss_handle *h, int level, int32_t *height, int32_t *width)
{
    if (!h || level < 0 || level > h->L) return ss_fail(SS_ERR_STATE, "bad level %d", level);
    if (height) *height = h->lh[level];
    if (width) *width = h->lw[level];
    return 0;
}

SS_EXPORT int ss_host_alloc(void **ptr, size_t bytes)
{
    SS_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return 0;
}
SS_EXPORT int ss_host_free(void *ptr)
{
    SS_CUDA(cudaFreeHost(ptr));
    return 0;
}
// 1 when ptr lies in page-locked host memory known to CUDA (asynchronous copies from it do not block), else 0
SS_EXPORT int ss_host_is_pinned(const void *ptr)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
    return a.type == cudaMemoryTypeHost ? 1 : 0;
}

// float-64 path: the chain of one frame (ss_f64.cuh).
static int ss64_launch_frame(ss_handle *h, int f)
{
    const int L = h->L;
    const size_t nb = (size_t)h->lh[L] * h->lw[L];
    const char *src = h->frames + (size_t)(f % h->n_slots) * h->frame_bytes + h->frame_shift;
    int rc;
    if ((rc = ss_timer_begin(h))) return rc;
    {
        dim3 g((h->cfg.width + 127) / 128, h->cfg.height);
        if (h->cfg.bytes_per_sample == 1)
            ss64_convert_kernel<uint8_t><<<g, 128, 0, h->s_comp>>>((const uint8_t *)src, (int)h->frame_pitch, h->cfg.height,
                                                                    h->cfg.width, h->g64[0]);
        else
            ss64_convert_kernel<uint16_t><<<g, 128, 0, h->s_comp>>>((const uint16_t *)src, (int)h->frame_pitch, h->cfg.height,
                                                                     h->cfg.width, h->g64[0]);
        h->launches++;
    }
    for (int l = 0; l < L; ++l) {
        dim3 g((h->lw[l + 1] + 127) / 128, h->lh[l + 1], 3);
        ss64_reduce_kernel<<<g, 128, 0, h->s_comp>>>(h->g64[l], h->lh[l], h->lw[l], h->g64[l + 1], h->lh[l + 1], h->lw[l + 1], h->k64);
        h->launches++;
    }
    for (int l = 0; l < L; ++l) {
        dim3 g((h->lw[l] + 127) / 128, h->lh[l]);
        ss64_lap_kernel<<<g, 128, 0, h->s_comp>>>(h->g64[l], h->lh[l], h->lw[l], h->g64[l + 1], h->lh[l + 1], h->lw[l + 1],
                                                   h->k64, h->lap64, h->sq64);
        ss64_energy_select_kernel<<<g, 128, 0, h->s_comp>>>(h->sq64, h->lh[l], h->lw[l], h->k64, h->lap64,
                                                             h->cfg.frame_index_base + f, h->have_state ? 0 : 1,
                                                             h->lv[l].best_e, h->lv[l].best_idx, h->fused64[l]);
        h->launches += 2;
    }
    SS_CUDA(cudaMemcpyAsync(h->base64 + (size_t)f * 3 * nb, h->g64[L], 3 * nb * sizeof(double), cudaMemcpyDeviceToDevice, h->s_comp));
    SS_CUDA(cudaGetLastError());
    if ((rc = ss_timer_end(h))) return rc;
    h->have_state = true;
    return 0;
}

// Launch the per-level kernel chain for frames [f0, f1) (all inside one batch slot).
static int ss_launch_batch(ss_handle *h, int f0, int f1)
{
    if (h->f64) {
        for (int f = f0; f < f1; ++f) {
            const int rc = ss64_launch_frame(h, f);
            if (rc) return rc;
        }
        return 0;
    }
    if (h->peer_on && !h->have_state && h->peer_epoch > 0) {
        // a new stack overwrites the running state: every rank must have finished reading the previous one
        ss_flag_wait_kernel<<<1, 32, 0, h->s_comp>>>(h->flags, h->peer_world, SS_FLAG_DONE, h->peer_epoch);
        h->launches++;
    }
    const int nf = f1 - f0;
    const int slot0 = f0 % h->n_slots;
    const size_t nb = (size_t)h->lh[h->L] * h->lw[h->L];
    int rc;
    if ((rc = ss_timer_begin(h))) return rc;
    if (h->L == 0) {
        for (int f = f0; f < f1; ++f) {
            const char *src = h->frames + (size_t)(f % h->n_slots) * h->frame_bytes + h->frame_shift;
            dim3 g((h->cfg.width + 255) / 256, h->cfg.height);
            if (h->cfg.bytes_per_sample == 1)
                ss_convert_kernel<uint8_t><<<g, 256, 0, h->s_comp>>>((const uint8_t *)src, (int)h->frame_pitch,
                    h->cfg.height, h->cfg.width, h->base + (size_t)f * 3 * nb);
            else
                ss_convert_kernel<uint16_t><<<g, 256, 0, h->s_comp>>>((const uint16_t *)src, (int)h->frame_pitch,
                    h->cfg.height, h->cfg.width, h->base + (size_t)f * 3 * nb);
            h->launches++;
        }
    }
    for (int l = 0; l < h->L; ++l) {
        SsLevel &v = h->lv[l];
        SsLevelParams p;
        memset(&p, 0, sizeof(p));
        p.h = v.h; p.w = v.w; p.h2 = h->lh[l + 1]; p.w2 = h->lw[l + 1];
        if (l == 0) {
            p.in = h->frames + (size_t)slot0 * h->frame_bytes + h->frame_shift;
            p.tma_row0 = slot0 * h->cfg.height;
            p.in_frame_stride = (long long)h->frame_bytes;
            p.in_pitch = (int)h->frame_pitch;
        } else {
            p.in = h->lv[l - 1].gnext;
            p.lap_scratch = h->lap_scratch;
            p.in_pitch = h->lv[l - 1].gpitch * 4;
            p.in_plane_stride = (long long)v.h * p.in_pitch;
            p.in_frame_stride = 3 * p.in_plane_stride;
        }
        if (l + 1 < h->L) { p.gnext = v.gnext; p.gnext_pitch = v.gpitch; }
        else { p.gnext = h->base + (size_t)f0 * 3 * nb; p.gnext_pitch = p.w2; }  // base level: tightly packed
        p.gnext_plane_stride = (long long)p.h2 * p.gnext_pitch;
        p.gnext_frame_stride = 3 * p.gnext_plane_stride;
        int groups = v.max_groups < nf ? v.max_groups : nf;
        if (groups < 1) groups = 1;
        const size_t npx = (size_t)v.h * v.w;
        if (groups > 1) {
            p.best_e = v.grp_e; p.best_idx = v.grp_idx; p.fused = v.grp_fused;
            p.group_stride = (long long)npx;
        } else {
            p.best_e = v.best_e; p.best_idx = v.best_idx; p.fused = v.fused;
        }
        p.groups = groups;
        p.n_frames = nf;
        p.frame_index0 = h->cfg.frame_index_base + f0;
        p.init = h->have_state ? 0 : 1;
        memcpy(p.k25, h->k25, sizeof(p.k25));
        p.skip_factor = getenv("SS_NO_SKIP") ? 0.f : ss_skip_factor(h->k25);
        dim3 grid((v.w + SS_TX - 1) / SS_TX, (v.h + SS_TY - 1) / SS_TY, groups);
        size_t smem = SS_SMEM_BYTES;
        if (const char *xs = getenv("SS_EXTRA_SMEM")) {   // experiment knob: e.g. 20000 forces one CTA per SM
            smem += (size_t)atoi(xs);
            cudaFuncSetAttribute(ss_level_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(ss_level_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(ss_level_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        if (h->profile) {
            if (h->ltimers_used == h->ltimers.size()) {
                ss_handle::LevelTimer t;
                SS_CUDA(cudaEventCreate(&t.a));
                SS_CUDA(cudaEventCreate(&t.b));
                h->ltimers.push_back(t);
            }
            h->ltimers[h->ltimers_used].level = l;
            SS_CUDA(cudaEventRecord(h->ltimers[h->ltimers_used].a, h->s_comp));
        }
        if (l == 0 && h->cfg.bytes_per_sample == 1)
            ss_level_kernel<uint8_t><<<grid, SS_NT, smem, h->s_comp>>>(p, h->tmap_frames);
        else if (l == 0)
            ss_level_kernel<uint16_t><<<grid, SS_NT, smem, h->s_comp>>>(p, h->tmap_frames);
        else
            ss_level_kernel<float><<<grid, SS_NT, smem, h->s_comp>>>(p, h->tmap_frames);
        h->launches++;
        if (groups > 1) {
            ss_merge_groups_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, h->s_comp>>>(
                npx, groups, v.grp_e, v.grp_idx, v.grp_fused, h->have_state ? 1 : 0, v.best_e, v.best_idx, v.fused);
            h->launches++;
        }
        if (h->profile) {
            SS_CUDA(cudaEventRecord(h->ltimers[h->ltimers_used].b, h->s_comp));
            h->ltimers_used++;
        }
    }
    SS_CUDA(cudaGetLastError());
    if ((rc = ss_timer_end(h))) return rc;
    h->have_state = true;
    return 0;
}

// Launch every complete batch (or, with flush, also the trailing partial one).
static int ss_launch_pending(ss_handle *h, bool flush)
{
    while (h->n_launched < h->n_pushed) {
        const int f0 = h->n_launched;
        int f1 = f0 + h->batch - (f0 % h->batch);
        if (f1 > h->n_pushed) {
            if (!flush) break;
            f1 = h->n_pushed;
        }
        const int bslot = (f0 % h->n_slots) / h->batch;
        SS_CUDA(cudaEventRecord(h->ev_copied[bslot], h->s_copy));
        SS_CUDA(cudaStreamWaitEvent(h->s_comp, h->ev_copied[bslot], 0));
        const int rc = ss_launch_batch(h, f0, f1);
        if (rc) return rc;
        SS_CUDA(cudaEventRecord(h->ev_done[bslot], h->s_comp));
        h->n_launched = f1;
    }
    return 0;
}

// R <-> B of every pixel of one stored frame, in place: 4 pixels (three 32-bit words) per thread for 8-bit samples,
// 2 pixels for 16-bit ones; rows of the frame store are 4-byte aligned.
template <typename T>
__global__ void ss_swap_rb_kernel(char *frame, size_t pitch, int w, int h)
{
    const int y = blockIdx.y;
    const int per = sizeof(T) == 1 ? 4 : 2;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * per;
    if (x >= w || y >= h) return;
    T *px = (T *)(frame + (size_t)y * pitch) + 3 * x;
    if (x + per <= w) {
        uint32_t *wd = (uint32_t *)px;
        const uint32_t a = wd[0], b = wd[1], c = wd[2];
        if (sizeof(T) == 1) {   // bytes  R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3  ->  B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
            wd[0] = ((a >> 16) & 0xffu) | (a & 0xff00u) | ((a & 0xffu) << 16) | ((b & 0xff00u) << 16);
            wd[1] = (b & 0xffu) | ((a >> 24) << 8) | ((c & 0xffu) << 16) | (b & 0xff000000u);
            wd[2] = ((b >> 16) & 0xffu) | ((c >> 24) << 8) | (c & 0xff0000u) | ((c & 0xff00u) << 16);
        } else {                // half-words  R0 G0 | B0 R1 | G1 B1  ->  B0 G0 | R0 B1 | G1 R1
            wd[0] = (b & 0xffffu) | (a & 0xffff0000u);
            wd[1] = (a & 0xffffu) | (c & 0xffff0000u);
            wd[2] = (c & 0xffffu) | (b & 0xffff0000u);
        }
    } else {
        for (int k = 0; x + k < w; ++k) { const T t = px[3 * k]; px[3 * k] = px[3 * k + 2]; px[3 * k + 2] = t; }
    }
}

static int ss_push_impl(ss_handle *h, const void *src, size_t pitch, cudaMemcpyKind kind, bool swap_rb = false)
{
    if (!h || !src) return ss_fail(SS_ERR_STATE, "null argument");
    if (h->finished_begin) return ss_fail(SS_ERR_STATE, "push after finish_begin; call ss_reset first");
    if (h->n_pushed >= h->cfg.max_frames)
        return ss_fail(SS_ERR_STATE, "more than max_frames=%d frames pushed", h->cfg.max_frames);
    const size_t row = (size_t)h->cfg.width * 3 * h->cfg.bytes_per_sample;
    if (pitch < row) return ss_fail(SS_ERR_SHAPE, "row pitch %zu smaller than a row (%zu bytes)", pitch, row);
    SS_CUDA(cudaSetDevice(h->cfg.device));
    const int slot = h->n_pushed % h->n_slots;
    if (h->n_pushed >= h->n_slots && slot % h->batch == 0) {
        // the ring wraps: the batch that used these slots must have been consumed
        SS_CUDA(cudaStreamWaitEvent(h->s_copy, h->ev_done[slot / h->batch], 0));
    }
    char *dst = h->frames + (size_t)slot * h->frame_bytes + h->frame_shift;
    if (pitch == row && h->frame_pitch == row)   // contiguous on both sides: one 1-D transfer
        SS_CUDA(cudaMemcpyAsync(dst, src, row * h->cfg.height, kind, h->s_copy));
    else
        SS_CUDA(cudaMemcpy2DAsync(dst, h->frame_pitch, src, pitch, row, h->cfg.height, kind, h->s_copy));
    if (swap_rb) {     // behind the copy, ahead of the batch's kernels (same stream as the copy)
        const int per = h->cfg.bytes_per_sample == 1 ? 4 : 2;
        dim3 g(((h->cfg.width + per - 1) / per + 127) / 128, h->cfg.height);
        if (h->cfg.bytes_per_sample == 1)
            ss_swap_rb_kernel<uint8_t><<<g, 128, 0, h->s_copy>>>(dst, h->frame_pitch, h->cfg.width, h->cfg.height);
        else
            ss_swap_rb_kernel<uint16_t><<<g, 128, 0, h->s_copy>>>(dst, h->frame_pitch, h->cfg.width, h->cfg.height);
        h->launches++;
        SS_CUDA(cudaGetLastError());
    }
    h->n_pushed++;
    if (h->n_pushed % h->batch == 0) return ss_launch_pending(h, false);
    return 0;
}

SS_EXPORT int ss_push_frame(ss_handle *h, const void *bgr, size_t row_pitch_bytes)
{
    return ss_push_impl(h, bgr, row_pitch_bytes, cudaMemcpyHostToDevice);
}

// Same as ss_push_frame for a frame whose samples are in R, G, B order (what an uncompressed TIFF holds): the
// frame is uploaded as it is and the channels are swapped on the device, so the host only moves bytes.
SS_EXPORT int ss_push_frame_rgb(ss_handle *h, const void *rgb, size_t row_pitch_bytes)
{
    return ss_push_impl(h, rgb, row_pitch_bytes, cudaMemcpyHostToDevice, true);
}
SS_EXPORT int ss_push_frame_device(ss_handle *h, const void *d_bgr, size_t row_pitch_bytes)
{
    return ss_push_impl(h, d_bgr, row_pitch_bytes, cudaMemcpyDeviceToDevice);
}

// Block until every pushed frame has reached the device (host staging buffers may be reused).
SS_EXPORT int ss_sync_copies(ss_handle *h)
{
    if (!h) return ss_fail(SS_ERR_STATE, "null handle");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    SS_CUDA(cudaStreamSynchronize(h->s_copy));
    return 0;
}

SS_EXPORT int ss_sync(ss_handle *h)
{
    if (!h) return ss_fail(SS_ERR_STATE, "null handle");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    SS_CUDA(cudaStreamSynchronize(h->s_copy));
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    return 0;
}

// Re-run the kernel chain over the frames that are still resident in the ring (benchmarks: inputs
// already in HBM).  Requires that no slot has been overwritten.
SS_EXPORT int ss_rerun_resident(ss_handle *h)
{
    if (!h) return ss_fail(SS_ERR_STATE, "null handle");
    if (h->n_pushed > h->n_slots)
        return ss_fail(SS_ERR_STATE, "%d frames pushed but only %d resident slots", h->n_pushed, h->n_slots);
    SS_CUDA(cudaSetDevice(h->cfg.device));
    SS_CUDA(cudaStreamSynchronize(h->s_copy));
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    h->n_launched = 0;
    h->finished_begin = false;
    h->have_state = false;
    h->timers_used = 0;
    h->ltimers_used = 0;
    h->launches = 0;
    return ss_launch_pending(h, true);
}

SS_EXPORT int ss_finish_begin(ss_handle *h, int64_t *counts)
{
    if (!h) return ss_fail(SS_ERR_STATE, "null handle");
    if (h->n_pushed == 0) return ss_fail(SS_ERR_STATE, "no frame pushed");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    int rc;
    if (!h->finished_begin) {
        if ((rc = ss_launch_pending(h, true))) return rc;
        const int nb = h->lh[h->L] * h->lw[h->L];
        if ((rc = ss_timer_begin(h))) return rc;
        SS_CUDA(cudaMemsetAsync(h->counts, 0, sizeof(unsigned int) * h->num_values * h->n_pushed, h->s_comp));
        if (h->f64) {
            for (int f = 0; f < h->n_pushed; ++f) {
                ss64_base_gray_hist_kernel<<<(nb + 255) / 256, 256, 0, h->s_comp>>>(h->base64 + (size_t)f * 3 * nb, nb,
                    h->gray + (size_t)f * nb, h->counts + (size_t)f * h->num_values, h->num_values);
                h->launches++;
            }
        } else {
            dim3 g((nb + 255) / 256, h->n_pushed);
            ss_base_gray_hist_kernel<<<g, 256, 0, h->s_comp>>>(h->base, nb, h->gray, h->counts, h->num_values);
            h->launches++;
        }
        SS_CUDA(cudaGetLastError());
        if ((rc = ss_timer_end(h))) return rc;
        h->finished_begin = true;
    }
    if (counts) {
        const size_t n = (size_t)h->n_pushed * h->num_values;
        std::vector<unsigned int> tmp(n);
        SS_CUDA(cudaMemcpyAsync(tmp.data(), h->counts, n * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->s_comp));
        SS_CUDA(cudaStreamSynchronize(h->s_comp));
        for (size_t i = 0; i < n; ++i) counts[i] = (int64_t)tmp[i];
    }
    return 0;
}

// Base-level selection on this rank's frames (entropy / deviation arg-max + winners' pixels).
static int ss_select_base_impl(ss_handle *h, const float *log_prob, int per_px)
{
    if (!h || !h->finished_begin) return ss_fail(SS_ERR_STATE, "ss_finish_begin must be called first");
    if (h->f64) return ss_fail(SS_ERR_STATE, "float-64 handle: use ss_finish_select_base_f64");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    const int nb = h->lh[h->L] * h->lw[h->L];
    const size_t n = (size_t)h->n_pushed * (per_px ? (size_t)nb : (size_t)h->num_values);
    std::vector<float> tab;
    if (!log_prob) {
        // C-only callers: logf from libm; numpy's SIMD log can differ from it in the last ulp, which
        // matters only for exact ties-by-rounding of the base-level entropy arg-max.
        const size_t nc = (size_t)h->n_pushed * h->num_values;
        std::vector<unsigned int> cnt(nc);
        SS_CUDA(cudaMemcpyAsync(cnt.data(), h->counts, nc * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->s_comp));
        SS_CUDA(cudaStreamSynchronize(h->s_comp));
        tab.resize(nc);
        for (int f = 0; f < h->n_pushed; ++f)
            for (int v = 0; v < h->num_values; ++v) {
                const unsigned int c = cnt[(size_t)f * h->num_values + v];
                const float pr = (float)((double)(float)c / (double)nb);
                tab[(size_t)f * h->num_values + v] = c ? logf(pr) : -INFINITY;
            }
        log_prob = tab.data();
        per_px = 0;
    }
    SS_CUDA(cudaMemcpyAsync(h->log_p, log_prob, n * sizeof(float), cudaMemcpyHostToDevice, h->s_comp));
    if (!tab.empty()) SS_CUDA(cudaStreamSynchronize(h->s_comp));
    int rc;
    if ((rc = ss_timer_begin(h))) return rc;
    {
        dim3 g((nb + 63) / 64, h->n_pushed);
        if (h->side <= SS_SMALL_WINDOW)
            ss_base_metrics_kernel<SS_SMALL_WINDOW><<<g, 64, 0, h->s_comp>>>(h->gray, h->log_p, per_px, h->lh[h->L], h->lw[h->L],
                                                                             h->num_values, h->side, h->ent, h->dev);
        else
            ss_base_metrics_kernel<SS_MAX_WINDOW><<<g, 64, 0, h->s_comp>>>(h->gray, h->log_p, per_px, h->lh[h->L], h->lw[h->L],
                                                                           h->num_values, h->side, h->ent, h->dev);
        h->launches++;
    }
    ss_base_select_kernel<<<(nb + 127) / 128, 128, 0, h->s_comp>>>(h->ent, h->dev, h->base, nb, h->n_pushed,
        h->cfg.frame_index_base, h->best_ve, h->best_ie, h->best_vd, h->best_id, h->pix_e, h->pix_d);
    h->launches++;
    SS_CUDA(cudaGetLastError());
    if ((rc = ss_timer_end(h))) return rc;
    return 0;
}

SS_EXPORT int ss_finish_select_base(ss_handle *h, const float *log_prob) { return ss_select_base_impl(h, log_prob, 0); }

// The table log(float32(c) / n_base) for c = 0 .. n_base (n = n_base + 1 entries), as the caller's log computes it.
SS_EXPORT int ss_set_log_lut(ss_handle *h, const float *lut, int n)
{
    if (!h || !lut) return ss_fail(SS_ERR_STATE, "null argument");
    const int nb = h->lh[h->L] * h->lw[h->L];
    if (n != nb + 1) return ss_fail(SS_ERR_SHAPE, "log table must have %d entries (base level %d x %d), got %d", nb + 1,
                                    h->lh[h->L], h->lw[h->L], n);
    SS_CUDA(cudaSetDevice(h->cfg.device));
    SS_CUDA(cudaMemcpyAsync(h->log_lut, lut, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, h->s_comp));
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    return 0;
}

// Histogram + base-level selection entirely on the device (no host round trip, no synchronisation): flushes
// pending frames, then gray / histogram, window entropy through the count -> log-probability table, deviation,
// arg-max and winners' pixels.  float-32 handles only.
SS_EXPORT int ss_finish_base_device(ss_handle *h)
{
    if (!h) return ss_fail(SS_ERR_STATE, "null handle");
    if (h->f64) return ss_fail(SS_ERR_STATE, "float-64 handle: use ss_finish_begin + ss_finish_select_base_f64");
    int rc = ss_finish_begin(h, nullptr);
    if (rc) return rc;
    const int nb = h->lh[h->L] * h->lw[h->L];
    if ((rc = ss_timer_begin(h))) return rc;
    dim3 g((nb + 63) / 64, h->n_pushed);
    if (h->side <= SS_SMALL_WINDOW)
        ss_base_metrics_kernel<SS_SMALL_WINDOW><<<g, 64, 0, h->s_comp>>>(h->gray, h->log_lut, 2, h->lh[h->L], h->lw[h->L],
                                                                         h->num_values, h->side, h->ent, h->dev, h->counts);
    else
        ss_base_metrics_kernel<SS_MAX_WINDOW><<<g, 64, 0, h->s_comp>>>(h->gray, h->log_lut, 2, h->lh[h->L], h->lw[h->L],
                                                                       h->num_values, h->side, h->ent, h->dev, h->counts);
    ss_base_select_kernel<<<(nb + 127) / 128, 128, 0, h->s_comp>>>(h->ent, h->dev, h->base, nb, h->n_pushed,
        h->cfg.frame_index_base, h->best_ve, h->best_ie, h->best_vd, h->best_id, h->pix_e, h->pix_d);
    h->launches += 2;
    SS_CUDA(cudaGetLastError());
    return ss_timer_end(h);
}

// The fused image of the last ss_finish_collapse(h, NULL, 0), device -> host.
SS_EXPORT int ss_download_output(ss_handle *h, void *out_bgr, size_t out_row_pitch_bytes)
{
    if (!h || !out_bgr) return ss_fail(SS_ERR_STATE, "null argument");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    const size_t row = (size_t)h->cfg.width * 3 * h->cfg.bytes_per_sample;
    if (out_row_pitch_bytes < row) return ss_fail(SS_ERR_SHAPE, "output pitch too small");
    SS_CUDA(cudaMemcpy2DAsync(out_bgr, out_row_pitch_bytes, h->out_dev, h->out_pitch, row, h->cfg.height,
                              cudaMemcpyDeviceToHost, h->s_comp));
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    return 0;
}

// Same, with the table given per base pixel: log_prob_px[f][i] = log(probability of the gray value of pixel i
// of frame f).  Cheap for 16-bit stacks, where a 65536-entry table per frame is mostly empty.
SS_EXPORT int ss_finish_select_base_px(ss_handle *h, const float *log_prob_px)
{
    if (!log_prob_px) return ss_fail(SS_ERR_STATE, "null table");
    return ss_select_base_impl(h, log_prob_px, 1);
}

// float_type = 'float-64': the table np.log(probabilities) in float64 (n_frames x num_values, or per base pixel
// when per_px != 0); NULL -> libm log.  Also blends the base (the float-64 state is not exchangeable).
SS_EXPORT int ss_finish_select_base_f64(ss_handle *h, const double *log_prob, int per_px)
{
    if (!h || !h->finished_begin) return ss_fail(SS_ERR_STATE, "ss_finish_begin must be called first");
    if (!h->f64) return ss_fail(SS_ERR_STATE, "not a float-64 handle");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    const int nb = h->lh[h->L] * h->lw[h->L];
    std::vector<double> tab;
    if (!log_prob) {
        const size_t nc = (size_t)h->n_pushed * h->num_values;
        std::vector<unsigned int> cnt(nc);
        SS_CUDA(cudaMemcpyAsync(cnt.data(), h->counts, nc * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->s_comp));
        SS_CUDA(cudaStreamSynchronize(h->s_comp));
        tab.resize(nc);
        for (size_t i = 0; i < nc; ++i) tab[i] = cnt[i] ? log((double)cnt[i] / (double)nb) : -INFINITY;
        log_prob = tab.data();
        per_px = 0;
    }
    const size_t n = (size_t)h->n_pushed * (per_px ? (size_t)nb : (size_t)h->num_values);
    SS_CUDA(cudaMemcpyAsync(h->logp64, log_prob, n * sizeof(double), cudaMemcpyHostToDevice, h->s_comp));
    if (!tab.empty()) SS_CUDA(cudaStreamSynchronize(h->s_comp));
    int rc;
    if ((rc = ss_timer_begin(h))) return rc;
    dim3 g((nb + 63) / 64, h->n_pushed);
    if (h->side <= SS_SMALL_WINDOW)
        ss64_base_metrics_kernel<SS_SMALL_WINDOW><<<g, 64, 0, h->s_comp>>>(h->gray, h->logp64, per_px, h->lh[h->L], h->lw[h->L],
                                                                           h->num_values, h->side, h->ent64, h->dev64);
    else
        ss64_base_metrics_kernel<SS_MAX_WINDOW><<<g, 64, 0, h->s_comp>>>(h->gray, h->logp64, per_px, h->lh[h->L], h->lw[h->L],
                                                                         h->num_values, h->side, h->ent64, h->dev64);
    ss64_base_select_fuse_kernel<<<(nb + 127) / 128, 128, 0, h->s_comp>>>(h->ent64, h->dev64, h->base64, nb, h->n_pushed,
        h->cfg.frame_index_base, h->best_ie, h->best_id, h->fusedbase64);
    h->launches += 2;
    SS_CUDA(cudaGetLastError());
    if ((rc = ss_timer_end(h))) return rc;
    return 0;
}

// Truncated gray base level of every pushed frame (n_frames x hL x wL int32), after ss_finish_begin.
SS_EXPORT int ss_get_base_gray(ss_handle *h, int32_t *out)
{
    if (!h || !out || !h->finished_begin) return ss_fail(SS_ERR_STATE, "ss_finish_begin must be called first");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->n_pushed * h->lh[h->L] * h->lw[h->L];
    SS_CUDA(cudaMemcpyAsync(out, h->gray, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->s_comp));
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    return 0;
}

// Blend the base, collapse the fused pyramid, convert and download (reference pyramid.py:106-111,
// 57-64, 179).
SS_EXPORT int ss_finish_collapse(ss_handle *h, void *out_bgr, size_t out_row_pitch_bytes)
{
    if (!h || !h->finished_begin) return ss_fail(SS_ERR_STATE, "ss_finish_begin must be called first");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    const int L = h->L;
    const size_t nb = (size_t)h->lh[L] * h->lw[L];
    int rc;
    if (h->f64) {
        if ((rc = ss_timer_begin(h))) return rc;
        const double *img = h->fusedbase64;
        if (L == 0) {
            dim3 g((h->cfg.width + 127) / 128, h->cfg.height);
            ss64_finalize_kernel<<<g, 128, 0, h->s_comp>>>(img, h->cfg.height, h->cfg.width, h->cfg.bytes_per_sample,
                                                            h->out_dev, (int)h->out_pitch);
            h->launches++;
        }
        for (int l = L - 1; l >= 0; --l) {
            double *dst = l == 0 ? nullptr : ((img == h->tmp64_a) ? h->tmp64_b : h->tmp64_a);
            dim3 g((h->lw[l] + 127) / 128, h->lh[l]);
            ss64_collapse_kernel<<<g, 128, 0, h->s_comp>>>(img, h->lh[l + 1], h->lw[l + 1], h->fused64[l], h->lh[l], h->lw[l],
                                                            h->k64, dst, h->out_dev, (int)h->out_pitch,
                                                            l == 0 ? h->cfg.bytes_per_sample : 0);
            h->launches++;
            img = dst;
        }
        SS_CUDA(cudaGetLastError());
        if ((rc = ss_timer_end(h))) return rc;
        if (out_bgr) {
            const size_t row = (size_t)h->cfg.width * 3 * h->cfg.bytes_per_sample;
            if (out_row_pitch_bytes < row) return ss_fail(SS_ERR_SHAPE, "output pitch too small");
            SS_CUDA(cudaMemcpy2DAsync(out_bgr, out_row_pitch_bytes, h->out_dev, h->out_pitch, row, h->cfg.height,
                                      cudaMemcpyDeviceToHost, h->s_comp));
        }
        SS_CUDA(cudaStreamSynchronize(h->s_comp));
        return 0;
    }
    if ((rc = ss_timer_begin(h))) return rc;
    ss_base_fuse_kernel<<<(unsigned)((3 * nb + 255) / 256), 256, 0, h->s_comp>>>(h->pix_e, h->pix_d, 3 * nb, h->fused_base);
    h->launches++;
    const float *img = h->fused_base;
    if (L == 0) {
        dim3 g((h->cfg.width + 255) / 256, h->cfg.height);
        ss_finalize_kernel<<<g, 256, 0, h->s_comp>>>(img, h->cfg.height, h->cfg.width, h->cfg.bytes_per_sample,
                                                      h->out_dev, (int)h->out_pitch);
        h->launches++;
    }
    for (int l = L - 1; l >= 0; --l) {
        SsCollapseParams p;
        memset(&p, 0, sizeof(p));
        p.img = img;
        p.layer = h->lv[l].fused;
        p.h = h->lh[l]; p.w = h->lw[l]; p.h2 = h->lh[l + 1]; p.w2 = h->lw[l + 1];
        memcpy(p.k25, h->k25, sizeof(p.k25));
        if (l == 0) {
            p.final_bytes = h->cfg.bytes_per_sample;
            p.out_px = h->out_dev;
            p.out_pitch = (int)h->out_pitch;
        } else {
            p.out = (img == h->tmp_a) ? h->tmp_b : h->tmp_a;
        }
        dim3 g((p.w + 63) / 64, (p.h + 15) / 16);
        ss_collapse_kernel<<<g, 256, 0, h->s_comp>>>(p);
        h->launches++;
        img = p.out;
    }
    SS_CUDA(cudaGetLastError());
    if ((rc = ss_timer_end(h))) return rc;
    if (out_bgr) {
        const size_t row = (size_t)h->cfg.width * 3 * h->cfg.bytes_per_sample;
        if (out_row_pitch_bytes < row) return ss_fail(SS_ERR_SHAPE, "output pitch too small");
        SS_CUDA(cudaMemcpy2DAsync(out_bgr, out_row_pitch_bytes, h->out_dev, h->out_pitch, row, h->cfg.height,
                                  cudaMemcpyDeviceToHost, h->s_comp));
    }
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    return 0;
}

SS_EXPORT int ss_finish_end(ss_handle *h, const float *log_prob, void *out_bgr, size_t out_row_pitch_bytes)
{
    if (h && h->f64 && log_prob) return ss_fail(SS_ERR_STATE, "float-64 handle: use ss_finish_select_base_f64");
    int rc = (h && h->f64) ? ss_finish_select_base_f64(h, nullptr, 0) : ss_finish_select_base(h, log_prob);
    if (rc) return rc;
    return ss_finish_collapse(h, out_bgr, out_row_pitch_bytes);
}

SS_EXPORT int ss_finish(ss_handle *h, void *out_bgr, size_t out_row_pitch_bytes)
{
    if (h && !h->f64) {
        const int rc = ss_finish_base_device(h);
        if (rc) return rc;
        return ss_finish_collapse(h, out_bgr, out_row_pitch_bytes);
    }
    int rc = ss_finish_begin(h, nullptr);
    if (rc) return rc;
    return ss_finish_end(h, nullptr, out_bgr, out_row_pitch_bytes);
}

SS_EXPORT int ss_get_argmax(ss_handle *h, int level, int32_t *out)
{
    if (!h || !out || level < 0 || level > h->L + 1) return ss_fail(SS_ERR_STATE, "bad argument");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    const int32_t *src;
    size_t n;
    if (level < h->L) { src = h->lv[level].best_idx; n = (size_t)h->lh[level] * h->lw[level]; }
    else { src = level == h->L ? h->best_ie : h->best_id; n = (size_t)h->lh[h->L] * h->lw[h->L]; }
    if (h->peer_on && h->peer_epoch > 0 && h->peer_world > 1) {
        // after a peer exchange the merged maps of the whole stack live on the root
        if (h->peer_rank != h->peer_root) return ss_fail(SS_ERR_STATE, "merged arg-max maps are on the root rank");
        src = h->mrg_idx + h->seg.px_off[level];
    }
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    SS_CUDA(cudaMemcpy(out, src, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

SS_EXPORT int ss_get_best_energy(ss_handle *h, int level, float *out)
{
    if (!h || !out || level < 0 || level >= h->L) return ss_fail(SS_ERR_STATE, "bad argument");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    SS_CUDA(cudaMemcpy(out, h->lv[level].best_e, (size_t)h->lh[level] * h->lw[level] * sizeof(float),
                       cudaMemcpyDeviceToHost));
    return 0;
}

SS_EXPORT int ss_get_fused_level(ss_handle *h, int level, float *out)
{
    if (!h || !out || level < 0 || level > h->L) return ss_fail(SS_ERR_STATE, "bad argument");
    if (h->f64) return ss_fail(SS_ERR_STATE, "not available on a float-64 handle");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->lh[level] * h->lw[level];
    const float *src = level < h->L ? h->lv[level].fused : h->fused_base;
    std::vector<float> tmp(3 * n);
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    SS_CUDA(cudaMemcpy(tmp.data(), src, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) out[3 * i + c] = tmp[c * n + i];
    return 0;
}

// g_{l+1} of frame `frame` (index within the most recent batch for l+1 < L; absolute for the base
// level l+1 == L) as 3 float32 planes -- test hook for reduce_layer (reference pyramid.py:27-32).
SS_EXPORT int ss_get_gauss_level(ss_handle *h, int level, int frame, float *out_planes)
{
    if (!h || !out_planes || level < 1 || level > h->L) return ss_fail(SS_ERR_STATE, "bad argument");
    if (h->f64) return ss_fail(SS_ERR_STATE, "not available on a float-64 handle");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    const size_t rows = 3 * (size_t)h->lh[level], wbytes = (size_t)h->lw[level] * sizeof(float);
    const size_t pitch = level < h->L ? (size_t)h->lv[level - 1].gpitch : (size_t)h->lw[level];
    const float *src = (level < h->L ? h->lv[level - 1].gnext : h->base) + (size_t)frame * rows * pitch;
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    SS_CUDA(cudaMemcpy2D(out_planes, wbytes, src, pitch * sizeof(float), wbytes, rows, cudaMemcpyDeviceToHost));
    return 0;
}

SS_EXPORT int ss_kernel_time_ms(ss_handle *h, float *ms)
{
    if (!h || !ms) return ss_fail(SS_ERR_STATE, "null argument");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    float total = 0.f;
    for (size_t i = 0; i < h->timers_used; ++i) {
        float t = 0.f;
        SS_CUDA(cudaEventElapsedTime(&t, h->timers[i].first, h->timers[i].second));
        total += t;
    }
    *ms = total;
    return 0;
}

// Time of timed segment i (batches first, then base histogram, base select, collapse).
SS_EXPORT int ss_segment_time_ms(ss_handle *h, int i, float *ms)
{
    if (!h || !ms || i < 0 || (size_t)i >= h->timers_used) return ss_fail(SS_ERR_STATE, "bad segment");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    SS_CUDA(cudaEventElapsedTime(ms, h->timers[i].first, h->timers[i].second));
    return 0;
}

// Per-launch CUDA-event timing of the fused level kernels (bench.py's roofline figure).
SS_EXPORT int ss_set_profiling(ss_handle *h, int on)
{
    if (!h) return ss_fail(SS_ERR_STATE, "null handle");
    h->profile = on != 0;
    return 0;
}

SS_EXPORT int ss_level_kernel_time_ms(ss_handle *h, int level, float *total_ms, int32_t *launches)
{
    if (!h || !total_ms || !launches) return ss_fail(SS_ERR_STATE, "null argument");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    float tot = 0.f;
    int n = 0;
    for (size_t i = 0; i < h->ltimers_used; ++i) {
        if (h->ltimers[i].level != level) continue;
        float t = 0.f;
        SS_CUDA(cudaEventElapsedTime(&t, h->ltimers[i].a, h->ltimers[i].b));
        tot += t;
        ++n;
    }
    *total_ms = tot;
    *launches = n;
    return 0;
}

SS_EXPORT int ss_launch_count(ss_handle *h, int64_t *launches)
{
    if (!h || !launches) return ss_fail(SS_ERR_STATE, "null argument");
    *launches = h->launches;
    return 0;
}

// ---- fused peer exchange: host side ------------------------------------------------------------------
#define SS_PEER_SLOTS 6   // IPC-shared allocations per rank: state_val, state_idx, state_coef, flags, out_dev, mrg_idx
static int ss_peer_alloc(ss_handle *h)
{
    if (h->f64) return ss_fail(SS_ERR_STATE, "the multi-GPU exchange is implemented for float-32 only");
    int rc;
    if (!h->flags) {
        if ((rc = ss_dalloc(&h->flags, 3 * SS_MAX_RANKS))) return rc;
        SS_CUDA(cudaMemset(h->flags, 0, 3 * SS_MAX_RANKS * sizeof(unsigned int)));
    }
    if (!h->mrg_coef && (rc = ss_dalloc(&h->mrg_coef, h->n_coeffs))) return rc;
    if (!h->mrg_idx && (rc = ss_dalloc(&h->mrg_idx, h->n_keys))) return rc;
    if (!h->d_peer_flags && (rc = ss_dalloc(&h->d_peer_flags, SS_MAX_RANKS))) return rc;
    return 0;
}

static void *const *ss_peer_own_ptrs(ss_handle *h, void **out)
{
    out[0] = h->state_val; out[1] = h->state_idx; out[2] = h->state_coef;
    out[3] = h->flags; out[4] = h->out_dev; out[5] = h->mrg_idx;
    return out;
}

// 6 CUDA IPC handles (64 bytes each) of this rank's exchangeable allocations.
SS_EXPORT int ss_peer_export(ss_handle *h, void *blob384)
{
    if (!h || !blob384) return ss_fail(SS_ERR_STATE, "null argument");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    int rc = ss_peer_alloc(h);
    if (rc) return rc;
    void *own[SS_PEER_SLOTS];
    ss_peer_own_ptrs(h, own);
    for (int i = 0; i < SS_PEER_SLOTS; ++i) {
        cudaIpcMemHandle_t mh;
        SS_CUDA(cudaIpcGetMemHandle(&mh, own[i]));
        memcpy((char *)blob384 + 64 * i, &mh, 64);
    }
    return 0;
}

// Row bands: the base level is cut into `world` bands of rows; the band of a finer level is twice the band of the
// next coarser one (so a band collapses into its own rows).  `need` = the rows of each level a rank must merge to
// collapse its band: its band, rounded to row pairs, plus the rows its 3-row expand windows reach.
static void ss_band_rows(const int *lh, int L, int world, int rank, int *row0, int *row1, int *own0, int *own1)
{
    std::vector<int> lo(L + 1), hi(L + 1);
    lo[L] = (int)((long long)rank * lh[L] / world);
    hi[L] = (int)((long long)(rank + 1) * lh[L] / world);
    for (int l = L - 1; l >= 0; --l) {
        lo[l] = 2 * lo[l + 1] < lh[l] ? 2 * lo[l + 1] : lh[l];
        hi[l] = rank == world - 1 ? lh[l] : (2 * hi[l + 1] < lh[l] ? 2 * hi[l + 1] : lh[l]);
    }
    int a = L > 0 ? lo[0] : 0, e = L > 0 ? hi[0] : 0;
    for (int l = 0; l < L; ++l) {
        own0[l] = lo[l]; own1[l] = hi[l];
        a = a / 2 * 2;                       // the collapse works on row pairs
        row0[l] = a; row1[l] = e;
        if (e > a) {                         // rows of level l+1 that the quads a/2 .. (e-1)/2 read
            const int na = a / 2 - 1 > 0 ? a / 2 - 1 : 0;
            int ne = (e - 1) / 2 + 2;
            if (ne > lh[l + 1]) ne = lh[l + 1];
            a = na; e = ne;
        }
    }
    for (int s2 = L; s2 < L + 2; ++s2) {     // base planes: every rank merges them whole (a few thousand pixels)
        own0[s2] = lo[L]; own1[s2] = hi[L];
        row0[s2] = 0; row1[s2] = lh[L];
    }
}

static void ss_peer_bands(ss_handle *h, int world, int rank)
{
    const int L = h->L;
    SsBandRows &b = h->band;
    memset(&b, 0, sizeof(b));
    b.count = L + 2;
    ss_band_rows(h->lh.data(), L, world, rank, b.row0, b.row1, b.own0, b.own1);
    for (int s2 = 0; s2 < L + 2; ++s2) b.width[s2] = h->lw[s2 < L ? s2 : L];
}

// Test hook (no GPU needed): the row bands of ss_peer_connect for an image of the given size; out = 4 x (L + 2) int32
// (row0 | row1 | own0 | own1 per segment), returns L.
SS_EXPORT int ss_test_peer_bands(int height, int width, int min_size, int world, int rank, int32_t *out, int32_t *level_h)
{
    std::vector<int> lh, lw;
    const int L = ss_levels_for(height, width, min_size, lh, lw);
    if (L + 2 > SS_MAX_SEGMENTS || !out || world < 1 || rank < 0 || rank >= world) return SS_ERR_STATE;
    int r[4][SS_MAX_SEGMENTS];
    ss_band_rows(lh.data(), L, world, rank, r[0], r[1], r[2], r[3]);
    for (int k = 0; k < 4; ++k)
        for (int s2 = 0; s2 < L + 2; ++s2) out[k * (L + 2) + s2] = r[k][s2];
    if (level_h)
        for (int l = 0; l <= L; ++l) level_h[l] = lh[l];
    return L;
}

static int ss_peer_setup(ss_handle *h, void *const (*ptrs)[SS_PEER_SLOTS], int world, int rank, int root)
{
    h->peer_world = world; h->peer_rank = rank; h->peer_root = root;
    memset(&h->pstate, 0, sizeof(h->pstate));
    h->pstate.world = world; h->pstate.rank = rank;
    unsigned int *pf[SS_MAX_RANKS] = {};
    for (int r = 0; r < world; ++r) {
        h->pstate.val[r] = (const float *)ptrs[r][0];
        h->pstate.idx[r] = (const int32_t *)ptrs[r][1];
        h->pstate.coef[r] = (const float *)ptrs[r][2];
        pf[r] = (unsigned int *)ptrs[r][3];
    }
    h->root_out = ptrs[root][4];
    h->root_mrg_idx = (int32_t *)ptrs[root][5];
    SS_CUDA(cudaMemcpy(h->d_peer_flags, pf, sizeof(pf), cudaMemcpyHostToDevice));
    ss_peer_bands(h, world, rank);
    h->peer_on = true;
    return 0;
}

// blobs = world x 384 bytes (ss_peer_export of every rank, in rank order; the own entry is ignored).
SS_EXPORT int ss_peer_connect(ss_handle *h, const void *blobs, int world, int rank, int root)
{
    if (!h || !blobs) return ss_fail(SS_ERR_STATE, "null argument");
    if (world < 1 || world > SS_MAX_RANKS || rank < 0 || rank >= world || root < 0 || root >= world)
        return ss_fail(SS_ERR_COMM, "bad world / rank / root");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    int rc = ss_peer_alloc(h);
    if (rc) return rc;
    void *ptrs[SS_MAX_RANKS][SS_PEER_SLOTS];
    for (int r = 0; r < world; ++r) {
        if (r == rank) { ss_peer_own_ptrs(h, ptrs[r]); continue; }
        for (int i = 0; i < SS_PEER_SLOTS; ++i) {
            if (h->peer_open[r][i]) { cudaIpcCloseMemHandle(h->peer_open[r][i]); h->peer_open[r][i] = nullptr; }
            cudaIpcMemHandle_t mh;
            memcpy(&mh, (const char *)blobs + ((size_t)r * SS_PEER_SLOTS + i) * 64, 64);
            SS_CUDA(cudaIpcOpenMemHandle(&h->peer_open[r][i], mh, cudaIpcMemLazyEnablePeerAccess));
            ptrs[r][i] = h->peer_open[r][i];
        }
    }
    return ss_peer_setup(h, ptrs, world, rank, root);
}

// The same with device pointers that are directly addressable from this device (several handles in one process:
// tests): ptrs = world x 6 pointers in the order state_val, state_idx, state_coef, flags, out_dev, mrg_idx.
SS_EXPORT int ss_peer_connect_ptrs(ss_handle *h, void *const *ptrs, int world, int rank, int root)
{
    if (!h || !ptrs) return ss_fail(SS_ERR_STATE, "null argument");
    if (world < 1 || world > SS_MAX_RANKS || rank < 0 || rank >= world || root < 0 || root >= world)
        return ss_fail(SS_ERR_COMM, "bad world / rank / root");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    int rc = ss_peer_alloc(h);
    if (rc) return rc;
    void *p2[SS_MAX_RANKS][SS_PEER_SLOTS];
    for (int r = 0; r < world; ++r)
        for (int i = 0; i < SS_PEER_SLOTS; ++i) p2[r][i] = ptrs[r * SS_PEER_SLOTS + i];
    ss_peer_own_ptrs(h, p2[rank]);
    return ss_peer_setup(h, p2, world, rank, root);
}

// This rank's 6 exchangeable device pointers (for ss_peer_connect_ptrs of the other handles of the process).
SS_EXPORT int ss_peer_ptrs(ss_handle *h, void **ptrs6)
{
    if (!h || !ptrs6) return ss_fail(SS_ERR_STATE, "null argument");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    int rc = ss_peer_alloc(h);
    if (rc) return rc;
    ss_peer_own_ptrs(h, ptrs6);
    return 0;
}

// Enqueue this rank's part of the exchange (all ranks call it once per stack, after their last frame):
// base selection on the own frames, READY signal, merge of the own bands out of the peers' memory, DONE signal,
// collapse of the own band into the root's output image, OUT signal.  `wait` != 0: also wait for it -- on the root
// for every rank's band, followed by the download into out_bgr (may be NULL: the image stays on the device).
// phase 1: base selection on the own frames, READY signal; phase 2: wait for every rank's READY, merge, DONE signal,
// banded collapse into the root's image, OUT signal; phase 3 (root): wait for every rank's OUT.
static int ss_peer_phase1(ss_handle *h)
{
    int rc = ss_finish_base_device(h);
    if (rc) return rc;
    const unsigned int epoch = ++h->peer_epoch;
    if ((rc = ss_timer_begin(h))) return rc;
    ss_flag_signal_kernel<<<1, 32, 0, h->s_comp>>>(h->d_peer_flags, h->peer_world, SS_FLAG_READY, h->peer_rank, epoch, -1);
    h->launches++;
    SS_CUDA(cudaGetLastError());
    return ss_timer_end(h);
}

static int ss_peer_phase3(ss_handle *h)
{
    if (h->peer_rank != h->peer_root) return 0;
    int rc;
    if ((rc = ss_timer_begin(h))) return rc;
    ss_flag_wait_kernel<<<1, 32, 0, h->s_comp>>>(h->flags, h->peer_world, SS_FLAG_OUT, h->peer_epoch);
    h->launches++;
    SS_CUDA(cudaGetLastError());
    return ss_timer_end(h);
}

static int ss_peer_phase2(ss_handle *h)
{
    int rc;
    const int L = h->L, world = h->peer_world, rank = h->peer_rank;
    const unsigned int epoch = h->peer_epoch;
    if ((rc = ss_timer_begin(h))) return rc;
    size_t total = 0;
    for (int s2 = 0; s2 < h->band.count; ++s2) total += (size_t)(h->band.row1[s2] - h->band.row0[s2]) * h->band.width[s2];
    ss_flag_wait_kernel<<<1, 32, 0, h->s_comp>>>(h->flags, world, SS_FLAG_READY, epoch);
    ss_peer_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->s_comp>>>(h->pstate, h->seg, h->band, h->flags, epoch,
                                                                                  h->mrg_coef, h->root_mrg_idx);
    ss_flag_signal_kernel<<<1, 32, 0, h->s_comp>>>(h->d_peer_flags, world, SS_FLAG_DONE, rank, epoch, -1);
    h->launches += 3;
    // base blend + banded collapse on the merged coefficients
    const size_t nb = (size_t)h->lh[L] * h->lw[L];
    ss_base_fuse_kernel<<<(unsigned)((3 * nb + 255) / 256), 256, 0, h->s_comp>>>(h->mrg_coef + h->seg.coef_off[L],
        h->mrg_coef + h->seg.coef_off[L + 1], 3 * nb, h->fused_base);
    h->launches++;
    const float *img = h->fused_base;
    if (L == 0) {
        // no Laplacian level: the blended base is the image; the own rows go to the root
        const int y0 = h->band.own0[L], y1 = h->band.own1[L];
        if (y1 > y0) {
            dim3 g((h->cfg.width + 255) / 256, y1 - y0);
            ss_finalize_kernel<<<g, 256, 0, h->s_comp>>>(img, h->cfg.height, h->cfg.width, h->cfg.bytes_per_sample,
                                                          h->root_out, (int)h->out_pitch, y0);
            h->launches++;
        }
    }
    for (int l = L - 1; l >= 0; --l) {
        const int r0 = l == 0 ? h->band.own0[0] / 2 * 2 : h->band.row0[l], r1 = l == 0 ? h->band.own1[0] : h->band.row1[l];
        SsCollapseParams p;
        memset(&p, 0, sizeof(p));
        p.img = img;
        p.layer = h->mrg_coef + h->seg.coef_off[l];
        p.h = h->lh[l]; p.w = h->lw[l]; p.h2 = h->lh[l + 1]; p.w2 = h->lw[l + 1];
        memcpy(p.k25, h->k25, sizeof(p.k25));
        p.qy0 = r0 / 2; p.y_end = r1;
        if (l == 0) {
            p.final_bytes = h->cfg.bytes_per_sample;
            p.out_px = h->root_out;
            p.out_pitch = (int)h->out_pitch;
        } else {
            p.out = (img == h->tmp_a) ? h->tmp_b : h->tmp_a;
        }
        if (r1 > r0) {
            dim3 g((p.w + 63) / 64, (r1 - r0 + 15) / 16);
            ss_collapse_kernel<<<g, 256, 0, h->s_comp>>>(p);
            h->launches++;
        }
        img = p.out;
    }
    ss_flag_signal_kernel<<<1, 32, 0, h->s_comp>>>(h->d_peer_flags, world, SS_FLAG_OUT, rank, epoch, h->peer_root);
    h->launches++;
    SS_CUDA(cudaGetLastError());
    return ss_timer_end(h);
}

// One phase of the exchange alone (a process that drives several ranks on ONE device must submit phase p of every
// rank before phase p + 1 of any: kernels of different streams can share a hardware queue, and a waiting kernel
// submitted ahead of the signal it waits for would block that queue).
SS_EXPORT int ss_peer_phase(ss_handle *h, int phase)
{
    if (!h || !h->peer_on) return ss_fail(SS_ERR_STATE, "ss_peer_connect first");
    if (h->n_pushed == 0) return ss_fail(SS_ERR_STATE, "no frame pushed");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    if (phase == 1) return ss_peer_phase1(h);
    if (phase == 2) return ss_peer_phase2(h);
    if (phase == 3) return ss_peer_phase3(h);
    return ss_fail(SS_ERR_STATE, "phase must be 1, 2 or 3");
}

SS_EXPORT int ss_peer_finish_async(ss_handle *h)
{
    if (!h || !h->peer_on) return ss_fail(SS_ERR_STATE, "ss_peer_connect first");
    if (h->n_pushed == 0) return ss_fail(SS_ERR_STATE, "no frame pushed");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    int rc = ss_peer_phase1(h);
    if (!rc) rc = ss_peer_phase2(h);
    if (!rc) rc = ss_peer_phase3(h);
    return rc;
}

SS_EXPORT int ss_peer_wait(ss_handle *h, void *out_bgr, size_t out_row_pitch_bytes)
{
    if (!h || !h->peer_on) return ss_fail(SS_ERR_STATE, "ss_peer_connect first");
    SS_CUDA(cudaSetDevice(h->cfg.device));
    if (h->peer_rank == h->peer_root && out_bgr) {
        const size_t row = (size_t)h->cfg.width * 3 * h->cfg.bytes_per_sample;
        if (out_row_pitch_bytes < row) return ss_fail(SS_ERR_SHAPE, "output pitch too small");
        SS_CUDA(cudaMemcpy2DAsync(out_bgr, out_row_pitch_bytes, h->out_dev, h->out_pitch, row, h->cfg.height,
                                  cudaMemcpyDeviceToHost, h->s_comp));
    }
    SS_CUDA(cudaStreamSynchronize(h->s_comp));
    return 0;
}

SS_EXPORT int ss_peer_finish(ss_handle *h, void *out_bgr, size_t out_row_pitch_bytes)
{
    const int rc = ss_peer_finish_async(h);
    if (rc) return rc;
    return ss_peer_wait(h, out_bgr, out_row_pitch_bytes);
}

// ---- test hooks -----------------------------------------------------------------------------------
// One collapse step on host arrays (3 float32 planes each) through the product's collapse kernel.
SS_EXPORT int ss_test_collapse_step(const float *img_planes, int h2, int w2, const float *layer_planes,
                                    int h, int w, double gen_kernel, float *out_planes)
{
    if (ss_device_count() == 0) return ss_fail(SS_ERR_NO_DEVICE, "no sm_100 device");
    float *d_img = nullptr, *d_lay = nullptr, *d_out = nullptr;
    const size_t n2 = 3 * (size_t)h2 * w2, n = 3 * (size_t)h * w;
    SS_CUDA(cudaMalloc(&d_img, n2 * 4));
    SS_CUDA(cudaMalloc(&d_out, n * 4));
    SS_CUDA(cudaMemcpy(d_img, img_planes, n2 * 4, cudaMemcpyHostToDevice));
    if (layer_planes) {
        SS_CUDA(cudaMalloc(&d_lay, n * 4));
        SS_CUDA(cudaMemcpy(d_lay, layer_planes, n * 4, cudaMemcpyHostToDevice));
    }
    SsCollapseParams p;
    memset(&p, 0, sizeof(p));
    p.img = d_img; p.layer = d_lay; p.out = d_out;
    p.h = h; p.w = w; p.h2 = h2; p.w2 = w2;
    ss_make_kernel(gen_kernel, p.k25);
    dim3 g((w + 63) / 64, (h + 15) / 16);
    ss_collapse_kernel<<<g, 256>>>(p);
    SS_CUDA(cudaGetLastError());
    SS_CUDA(cudaMemcpy(out_planes, d_out, n * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_img); cudaFree(d_lay); cudaFree(d_out);
    return 0;
}

#include "ss_dm_lib.cuh"
#include "ss_al_lib.cuh"
#include "ss_sub_lib.cuh"
