// ss_f64.cuh -- kernels of the float_type = 'float-64' path (reference base_stack_algo.py:15-16).
//
// With float64 images cv2.filter2D accumulates the 25 taps sequentially with separate multiply and add in
// float64 (no FMA, no tail-column rule; probed against cv2 4.13.0, tests/test_oracle_vs_reference.py), using
// the float64 kernel outer(k,k) as given.  Gray, energy and the per-level arg-max stay float32 (reference
// pyramid.py:49: lap.astype(np.float32)); entropy / deviation / base blend / collapse are float64.
// This path is the reference's non-default option: it is kept simple (one frame at a time, state in global
// memory, one thread per output) rather than tuned like the float-32 kernel.
#pragma once
#include <stdint.h>

__device__ __forceinline__ int ss64_refl(int i, int n)
{
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

struct Ss64Kernel { double k[25]; float kf[25]; };

// img.astype(float64) into planes (pyramid.py:126)
template <typename InT>
__global__ void ss64_convert_kernel(const InT *in, int pitch_bytes, int h, int w, double *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const InT *s = (const InT *)((const char *)in + (size_t)y * pitch_bytes) + 3 * x;
    const size_t plane = (size_t)h * w, o = (size_t)y * w + x;
    out[o] = (double)s[0];
    out[plane + o] = (double)s[1];
    out[2 * plane + o] = (double)s[2];
}

// reduce_layer (pyramid.py:27-32): blockIdx.z = channel plane
__global__ void ss64_reduce_kernel(const double *src, int h, int w, double *dst, int h2, int w2,
                                   const __grid_constant__ Ss64Kernel K)
{
    const int xo = blockIdx.x * blockDim.x + threadIdx.x, yo = blockIdx.y;
    if (xo >= w2) return;
    const double *s = src + (size_t)blockIdx.z * h * w;
    double acc = 0.0;
    for (int dy = 0; dy < 5; ++dy) {
        const double *row = s + (size_t)ss64_refl(2 * yo + dy - 2, h) * w;
        for (int dx = 0; dx < 5; ++dx)
            acc = __dadd_rn(acc, __dmul_rn(row[ss64_refl(2 * xo + dx - 2, w)], K.k[dy * 5 + dx]));
    }
    dst[((size_t)blockIdx.z * h2 + yo) * w2 + xo] = acc;
}

// 4 * convolve(zero-stuffed g_next) at (y, x), skipping the stuffed zeros (exact), pyramid.py:34-41
__device__ __forceinline__ double ss64_expand_at(const double *g, int h2, int w2, int y, int x, const double *k)
{
    const int H = 2 * h2, W = 2 * w2;
    double acc = 0.0;
    for (int dy = 0; dy < 5; ++dy) {
        const int yy = ss64_refl(y + dy - 2, H);
        if (yy & 1) continue;
        const double *row = g + (size_t)(yy >> 1) * w2;
        for (int dx = 0; dx < 5; ++dx) {
            const int xx = ss64_refl(x + dx - 2, W);
            if (xx & 1) continue;
            acc = __dadd_rn(acc, __dmul_rn(row[xx >> 1], k[dy * 5 + dx]));
        }
    }
    return __dmul_rn(4.0, acc);
}

// lap = g - expand(g_next) (pyramid.py:133-138) for the three planes + gray^2 of lap.astype(float32)
// (pyramid.py:49-50)
__global__ void ss64_lap_kernel(const double *g, int h, int w, const double *gn, int h2, int w2,
                                const __grid_constant__ Ss64Kernel K, double *lap, float *sq)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t n = (size_t)h * w, n2 = (size_t)h2 * w2, o = (size_t)y * w + x;
    float lf[3];
    for (int ch = 0; ch < 3; ++ch) {
        const double l = __dsub_rn(g[ch * n + o], ss64_expand_at(gn + ch * n2, h2, w2, y, x, K.k));
        lap[ch * n + o] = l;
        lf[ch] = (float)l;
    }
    const float gray = __fmaf_rn(lf[2], 0.299f, __fmaf_rn(lf[0], 0.114f, __fmul_rn(lf[1], 0.587f)));
    sq[o] = __fmul_rn(gray, gray);
}

// energy = convolve(gray^2) in float32 (FMA / tail-column rule of the float32 filter) + running strict-'>'
// arg-max over the frames; winner coefficients 0 + lap in float64 (pyramid.py:50-54)
__global__ void ss64_energy_select_kernel(const float *sq, int h, int w, const __grid_constant__ Ss64Kernel K,
                                          const double *lap, int gidx, int first, float *best_e,
                                          int32_t *best_idx, double *fused)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const bool tail = x >= w - (w % 8);
    float acc = 0.f;
    for (int dy = 0; dy < 5; ++dy) {
        const float *row = sq + (size_t)ss64_refl(y + dy - 2, h) * w;
        for (int dx = 0; dx < 5; ++dx) {
            const float v = row[ss64_refl(x + dx - 2, w)], k = K.kf[dy * 5 + dx];
            acc = tail ? __fadd_rn(acc, __fmul_rn(v, k)) : __fmaf_rn(v, k, acc);
        }
    }
    const size_t n = (size_t)h * w, o = (size_t)y * w + x;
    if (first || acc > best_e[o]) {
        best_e[o] = acc;
        best_idx[o] = gidx;
        for (int ch = 0; ch < 3; ++ch) fused[ch * n + o] = __dadd_rn(0.0, lap[ch * n + o]);
    }
}

// base level: gray = cvtColor(base.astype(float32)).astype(dtype) + histogram (pyramid.py:99-101,67)
__global__ void ss64_base_gray_hist_kernel(const double *base, int n_px, int32_t *gray, unsigned int *counts,
                                           int num_values)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const float b = (float)base[i], g = (float)base[(size_t)n_px + i], r = (float)base[2 * (size_t)n_px + i];
    const float v = __fmaf_rn(r, 0.299f, __fmaf_rn(b, 0.114f, __fmul_rn(g, 0.587f)));
    int gv = (int)v;
    gv = gv < 0 ? 0 : (gv >= num_values ? num_values - 1 : gv);
    gray[i] = gv;
    atomicAdd(&counts[gv], 1u);
}

__device__ __forceinline__ double ss64_np_pairwise_leaf(const double *a, int n)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}
// numpy's pairwise add.reduce (see ss_np_pairwise_sum): runs above 128 values are halved, DEPTH levels deep
template <int DEPTH>
__device__ double ss64_np_pairwise_sum(const double *a, int n)
{
    if (n <= 128) return ss64_np_pairwise_leaf(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(ss64_np_pairwise_sum<DEPTH - 1>(a, n2), ss64_np_pairwise_sum<DEPTH - 1>(a + n2, n - n2));
}
template <>
__device__ double ss64_np_pairwise_sum<0>(const double *a, int n) { return ss64_np_pairwise_leaf(a, n); }

// entropy / deviation in float64 (pyramid.py:66-93), blockIdx.y = frame
template <int MAXSIDE>
__global__ void ss64_base_metrics_kernel(const int32_t *gray, const double *log_p, int per_px, int h, int w,
                                         int num_values, int side, double *ent_out, double *dev_out)
{
    constexpr int DEPTH = MAXSIDE * MAXSIDE <= 128 ? 0 : 3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    const int n_px = h * w;
    if (i >= n_px) return;
    const int y = i / w, x = i - y * w;
    const int pad = (side - 1) / 2, n = side * side;
    double buf[MAXSIDE * MAXSIDE];
    int vals[MAXSIDE * MAXSIDE];
    const int32_t *g = gray + (size_t)f * n_px;
    const double *lp = log_p + (size_t)f * (per_px ? n_px : num_values);
    int c = 0;
    long long s = 0;
    for (int dy = -pad; dy <= pad; ++dy) {
        const int yy = ss64_refl(y + dy, h);
        for (int dx = -pad; dx <= pad; ++dx) {
            const int xx = ss64_refl(x + dx, w);
            const int v = g[yy * w + xx];
            vals[c] = v;
            buf[c] = __dmul_rn((double)v, lp[per_px ? yy * w + xx : v]);
            s += v;
            ++c;
        }
    }
    ent_out[(size_t)f * n_px + i] = __dmul_rn(-1.0, ss64_np_pairwise_sum<DEPTH>(buf, n));
    const double mean = (double)s / (double)n;
    for (int k = 0; k < n; ++k) {
        const double d = __dsub_rn((double)vals[k], mean);
        buf[k] = __dmul_rn(d, d);
    }
    dev_out[(size_t)f * n_px + i] = __ddiv_rn(ss64_np_pairwise_sum<DEPTH>(buf, n), (double)n);
}

// arg-max over frames + blend ((0 + base[best_e]) + base[best_d]) / 2 (pyramid.py:102-111)
__global__ void ss64_base_select_fuse_kernel(const double *ent, const double *dev, const double *base, int n_px,
                                             int n_frames, int frame_index0, int32_t *best_ie, int32_t *best_id,
                                             double *fused)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    double ve = ent[i], vd = dev[i];
    int ie = 0, id = 0;
    for (int f = 1; f < n_frames; ++f) {
        const double e = ent[(size_t)f * n_px + i], d = dev[(size_t)f * n_px + i];
        if (e > ve) { ve = e; ie = f; }
        if (d > vd) { vd = d; id = f; }
    }
    best_ie[i] = frame_index0 + ie;
    best_id[i] = frame_index0 + id;
    for (int ch = 0; ch < 3; ++ch) {
        const double a = base[((size_t)ie * 3 + ch) * n_px + i], b = base[((size_t)id * 3 + ch) * n_px + i];
        fused[(size_t)ch * n_px + i] = __ddiv_rn(__dadd_rn(__dadd_rn(0.0, a), b), 2.0);
    }
}

// one collapse step (pyramid.py:59-64) and, on the last one, clip(abs()) + truncating cast (pyramid.py:179)
__global__ void ss64_collapse_kernel(const double *img, int h2, int w2, const double *layer, int h, int w,
                                     const __grid_constant__ Ss64Kernel K, double *out, void *out_px, int out_pitch,
                                     int final_bytes)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t n = (size_t)h * w, n2 = (size_t)h2 * w2, o = (size_t)y * w + x;
    for (int ch = 0; ch < 3; ++ch) {
        const double v = __dadd_rn(ss64_expand_at(img + ch * n2, h2, w2, y, x, K.k), layer[ch * n + o]);
        if (!final_bytes) {
            out[ch * n + o] = v;
        } else {
            const double maxv = final_bytes == 1 ? 255.0 : 65535.0;
            double a = fabs(v);
            a = a > maxv ? maxv : a;
            const unsigned int u = __double2uint_rz(a);
            if (final_bytes == 1) ((uint8_t *)((char *)out_px + (size_t)y * out_pitch))[3 * x + ch] = (uint8_t)u;
            else ((uint16_t *)((char *)out_px + (size_t)y * out_pitch))[3 * x + ch] = (uint16_t)u;
        }
    }
}

__global__ void ss64_finalize_kernel(const double *img, int h, int w, int bytes, void *out, int pitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const double maxv = bytes == 1 ? 255.0 : 65535.0;
    for (int ch = 0; ch < 3; ++ch) {
        double a = fabs(img[((size_t)ch * h + y) * w + x]);
        a = a > maxv ? maxv : a;
        const unsigned int u = __double2uint_rz(a);
        if (bytes == 1) ((uint8_t *)((char *)out + (size_t)y * pitch))[3 * x + ch] = (uint8_t)u;
        else ((uint16_t *)((char *)out + (size_t)y * pitch))[3 * x + ch] = (uint16_t)u;
    }
}
