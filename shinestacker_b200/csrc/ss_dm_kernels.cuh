// ss_dm_kernels.cuh -- device code of the depth-map focus stacker (sm_100a), the second stack algorithm
// behind the reference's stack-algo boundary (reference src/shinestacker/algorithms/depth_map.py:10-123).
//
// One body, two builds, like ss_kernels.cuh: every kernel is a small functor "one call = one output
// pixel (x, y, z)"; nvcc instantiates dm_px_kernel<Functor> over a 2-D/3-D grid (the product), and
// tests/emu/ runs the same functors in plain loops under g++ (-DSS_HOST_EMU) so border handling, tap
// order and the whole launch sequence can be checked on a CPU-only box.  The emulation is test
// infrastructure; the product never uses it.
//
// Arithmetic (what the reference's cv2 / numpy calls evaluate, oracle/depthmap_oracle.py):
//   * gray        cv2.cvtColor(BGR2GRAY) on uint8/uint16: (B*3735 + G*19235 + R*9798 + 2^14) >> 15
//   * energy      'laplacian': |Laplacian_k(GaussianBlur_b(gray))|, 'sobel': |Sobel_x| + |Sobel_y|;
//                 separable float32 filters in cv2's symmetric form k0*x0 + sum_j kj*(x[+j] +- x[-j]),
//                 rows first, REFLECT101
//   * normalise   energies / max(energies) (one global maximum over the whole stack)
//   * smooth      cv2.bilateralFilter(e, d, 25, 25): disc of radius d/2, weight = space * colour
//   * focus map   'average': e_i / sum_i e_i;  'max': softmax((e_i - max_i e_i) / temperature)
//   * blend       Laplacian pyramid (cv2.pyrDown / cv2.pyrUp, [1,4,6,4,1]) of every frame, each level
//                 multiplied by the Gaussian pyramid of its weight map, summed over frames, collapsed,
//                 clip(|x|) and truncated to the input type.
// Everything is float32 with explicit, unfused multiplies and adds (the library is built with
// -fmad=false); all stages are HBM-streaming stencils except the bilateral filter (FP32-issue bound).
#pragma once
#include <stdint.h>
#include "ss_kernels.cuh"

#ifdef SS_HOST_EMU
#define SS_MDEV inline            /* member functions of the pixel functors */
#else
#define SS_MDEV __device__ __forceinline__
#endif

#define DM_MAX_TAPS 16   /* centre + up to 15 one-sided taps: separable kernels up to 31 wide */

struct DmTaps {          // one-sided view of a symmetric / anti-symmetric 1-D kernel
    float k[DM_MAX_TAPS]; // k[0] = centre tap, k[j] = tap at distance j
    int r;               // radius
    int anti;            // 1: anti-symmetric (pairs are x[+j] - x[-j]), 0: symmetric (x[+j] + x[-j])
};

#ifdef SS_HOST_EMU
SS_DEV float ss_div(float a, float b) { return a / b; }
SS_DEV unsigned __float_as_uint_compat(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
SS_DEV float __uint_as_float_compat(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
#else
SS_DEV float ss_div(float a, float b) { return __fdiv_rn(a, b); }
SS_DEV unsigned __float_as_uint_compat(float f) { return __float_as_uint(f); }
SS_DEV float __uint_as_float_compat(unsigned u) { return __uint_as_float(u); }
#endif

SS_DEV int dm_reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// cv2's small symmetric-kernel form along one axis of a float plane
SS_DEV float dm_filter1d(const float *line, long long stride, int pos, int n, const DmTaps &t)
{
    float acc = ss_mul(line[(long long)pos * stride], t.k[0]);
    for (int j = 1; j <= t.r; ++j) {
        const float a = line[(long long)dm_reflect101(pos + j, n) * stride];
        const float b = line[(long long)dm_reflect101(pos - j, n) * stride];
        const float pair = t.anti ? ss_add(a, -b) : ss_add(a, b);
        acc = ss_add(acc, ss_mul(pair, t.k[j]));
    }
    return acc;
}

template <typename InT>
SS_DEV float dm_gray_at(const void *frame, long long pitch, int x, int y)
{
    const InT *s = (const InT *)((const char *)frame + (long long)y * pitch) + 3 * x;
    const uint32_t v = (uint32_t)s[0] * 3735u + (uint32_t)s[1] * 19235u + (uint32_t)s[2] * 9798u + 16384u;
    return (float)(v >> 15);
}

// ---- pass 1, per frame: gray -> (blur) -> derivative filters -> |.| energy + statistics -----------------
// A: row pass of the Gaussian blur straight from the integer BGR frame (gray computed per tap); with
//    r == 0 taps {1} it just writes the gray plane (Sobel energy needs no blur).
template <typename InT>
struct DmGrayRowFilter {
    const void *frame; long long pitch; int h, w; DmTaps t; float *out;
    SS_MDEV void operator()(int x, int y, int) const
    {
        float acc = ss_mul(dm_gray_at<InT>(frame, pitch, x, y), t.k[0]);
        for (int j = 1; j <= t.r; ++j) {
            const float a = dm_gray_at<InT>(frame, pitch, dm_reflect101(x + j, w), y);
            const float b = dm_gray_at<InT>(frame, pitch, dm_reflect101(x - j, w), y);
            acc = ss_add(acc, ss_mul(ss_add(a, b), t.k[j]));
        }
        out[(long long)y * w + x] = acc;
    }
};

// B: column pass of a separable filter
struct DmColFilter {
    const float *in; int h, w; DmTaps t; float *out;
    SS_MDEV void operator()(int x, int y, int) const
    {
        out[(long long)y * w + x] = dm_filter1d(in + x, w, y, h, t);
    }
};

// C: the two row passes of the derivative pair (Laplacian: kd | ks, Sobel: [-1,0,1] | [1,2,1])
struct DmRowPair {
    const float *in; int h, w; DmTaps ta, tb; float *out_a, *out_b;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const float *line = in + (long long)y * w;
        out_a[(long long)y * w + x] = dm_filter1d(line, 1, x, w, ta);
        out_b[(long long)y * w + x] = dm_filter1d(line, 1, x, w, tb);
    }
};

// per-frame minimum / maximum of the (non-negative) energies as IEEE bit patterns, kept in DM_STAT_SLOTS
// (min, max) pairs per frame: every CTA updates the slot picked by its block index, so the atomics of a launch
// spread over 64 L2 lines instead of hammering one; the host folds the slots (ss_dm_engine.h, finish_begin)
#define DM_STAT_SLOTS 64
SS_DEV void dm_stats_update(unsigned *stats, float e)
{
    const unsigned bits = __float_as_uint_compat(e);
#ifdef SS_HOST_EMU
    if (bits < stats[0]) stats[0] = bits;
    if (bits > stats[1]) stats[1] = bits;
#else
    unsigned *slot = stats + 2 * ((blockIdx.x + 7 * blockIdx.y) & (DM_STAT_SLOTS - 1));
    const unsigned m = __activemask();
    const unsigned lo = __reduce_min_sync(m, bits), hi = __reduce_max_sync(m, bits);
    // one lane per warp; the values are monotone, so a plain (L2) read first skips almost every atomic
    if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) {
        if (lo < __ldcg(slot)) atomicMin(slot, lo);
        if (hi > __ldcg(slot + 1)) atomicMax(slot + 1, hi);
    }
#endif
}

// D: column passes + combination.  Laplacian: |col_b(a) + col_a(b)| (cv2.Laplacian = d2/dx2 + d2/dy2,
//    depth_map.py:40); Sobel: |col_b(a)| + |col_a(b)| (depth_map.py:32-33).  Energy of -0 is +0.
struct DmColCombine {
    const float *in_a, *in_b; int h, w; DmTaps ta, tb; int sobel; float *out; unsigned *stats;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const float da = dm_filter1d(in_a + x, w, y, h, tb);   // rows filtered with ta, columns with tb
        const float db = dm_filter1d(in_b + x, w, y, h, ta);
        const float e = sobel ? ss_add(fabsf(da), fabsf(db)) : fabsf(ss_add(da, db));
        out[(long long)y * w + x] = e;
        dm_stats_update(stats, e);
    }
};

// ---- after the last frame ---------------------------------------------------------------------------------
// energies / max_energy over the whole stack (depth_map.py:88-90); z = frame
struct DmNormalize {
    float *e; int h, w; float gmax;
    SS_MDEV void operator()(int x, int y, int z) const
    {
        float *p = e + ((long long)z * h + y) * w + x;
        *p = ss_div(*p, gmax);
    }
};

// cv2.bilateralFilter(e, d, 25, 25) (smooth_energy, depth_map.py:43-51).  sw holds the (2r+1)^2 space
// weights, 0 outside the disc.  Normalised energies lie in [0, 1], so the colour exponent
// t = -(v - v0)^2 / (2*25^2) is in [-8e-4, 0] and exp(t) = 1 + t + t^2/2 to 1e-10 -- three full-rate FP32
// operations instead of cv2's 4096-bin interpolated table or a MUFU.  `flat` frames (max - min < FLT_EPSILON)
// are copied, as cv2 does.
struct DmBilateral {
    const float *in; int h, w; int r; const float *sw; float c1, c2; const unsigned *stats; float gmax;
    float *out;
    SS_MDEV void operator()(int x, int y, int z) const
    {
        const long long plane = (long long)h * w;
        const float *src = in + (long long)z * plane;
        const float v0 = src[(long long)y * w + x];
        float mn = __uint_as_float_compat(stats[2 * z]), mx = __uint_as_float_compat(stats[2 * z + 1]);
        if (gmax > 0.f) { mn = ss_div(mn, gmax); mx = ss_div(mx, gmax); }
        if (fabsf(mn - mx) < 1.1920929e-07f) { out[(long long)z * plane + (long long)y * w + x] = v0; return; }
        float num = 0.f, den = 0.f;
        const int d = 2 * r + 1;
        if (x >= r && x < w - r && y >= r && y < h - r) {
            const float *row = src + (long long)(y - r) * w + (x - r);
            for (int i = 0; i < d; ++i, row += w) {
#pragma unroll 5
                for (int j = 0; j < d; ++j) {
                    const float s = sw[i * d + j];
                    if (s == 0.f) continue;
                    const float v = row[j];
                    const float dv = ss_add(v, -v0);
                    const float q = ss_mul(dv, dv);
                    const float wgt = ss_mul(s, ss_fma(q, ss_fma(q, c2, c1), 1.f));
                    num = ss_fma(v, wgt, num);
                    den = ss_add(den, wgt);
                }
            }
        } else {
            for (int i = 0; i < d; ++i) {
                const float *row = src + (long long)dm_reflect101(y - r + i, h) * w;
                for (int j = 0; j < d; ++j) {
                    const float s = sw[i * d + j];
                    if (s == 0.f) continue;
                    const float v = row[dm_reflect101(x - r + j, w)];
                    const float dv = ss_add(v, -v0);
                    const float q = ss_mul(dv, dv);
                    const float wgt = ss_mul(s, ss_fma(q, ss_fma(q, c2, c1), 1.f));
                    num = ss_fma(v, wgt, num);
                    den = ss_add(den, wgt);
                }
            }
        }
        out[(long long)z * plane + (long long)y * w + x] = ss_div(num, den);
    }
};

// The default diameter (smooth_size = 15 -> radius 7) with everything a tap needs in the kernel-parameter
// constant bank: s, s*c1 and s*c2 by value, loops fully unrolled, taps outside the disc removed at compile
// time, so an interior tap costs one load and six FP32 instructions:
//   dv = v - v0;  q = dv*dv;  wgt = fma(q, fma(q, s*c2, s*c1), s);  num = fma(v, wgt, num);  den += wgt.
template <int R>
struct DmBilateralFixed {
    enum { D = 2 * R + 1 };
    const float *in; int h, w; const unsigned *stats; float gmax; float *out;
    float s0[D * D], s1[D * D], s2[D * D];
    SS_MDEV void tap(float v, float v0, int t, float &num, float &den) const
    {
        const float dv = ss_add(v, -v0);
        const float q = ss_mul(dv, dv);
        const float wgt = ss_fma(q, ss_fma(q, s2[t], s1[t]), s0[t]);
        num = ss_fma(v, wgt, num);
        den = ss_add(den, wgt);
    }
    // any pixel (image borders: REFLECT101)
    SS_MDEV void one(const float *src, int x, int y, float *dst) const
    {
        const float v0 = src[(long long)y * w + x];
        float num = 0.f, den = 0.f;
        for (int i = 0; i < D; ++i) {
            const float *row = src + (long long)dm_reflect101(y - R + i, h) * w;
            for (int j = 0; j < D; ++j) {
                if ((i - R) * (i - R) + (j - R) * (j - R) > R * R) continue;
                tap(row[dm_reflect101(x - R + j, w)], v0, i * D + j, num, den);
            }
        }
        dst[(long long)y * w + x] = ss_div(num, den);
    }
    // One call = the two horizontally adjacent pixels (2 xi, y) and (2 xi + 1, y): away from the borders every
    // loaded sample feeds the windows of both (16 columns x 15 rows loaded for 2 x 149 taps), and the
    // tap constants are fetched once for the pair.  The order of the taps of each pixel is unchanged.
    SS_MDEV void operator()(int xi, int y, int z) const
    {
        const long long plane = (long long)h * w;
        const float *src = in + (long long)z * plane;
        float *dst = out + (long long)z * plane;
        const int x = 2 * xi;
        float mn = __uint_as_float_compat(stats[2 * z]), mx = __uint_as_float_compat(stats[2 * z + 1]);
        if (gmax > 0.f) { mn = ss_div(mn, gmax); mx = ss_div(mx, gmax); }
        if (fabsf(mn - mx) < 1.1920929e-07f) {      // cv2 copies a flat plane
            dst[(long long)y * w + x] = src[(long long)y * w + x];
            if (x + 1 < w) dst[(long long)y * w + x + 1] = src[(long long)y * w + x + 1];
            return;
        }
        if (!(x >= R && x + 1 < w - R && y >= R && y < h - R)) {
            one(src, x, y, dst);
            if (x + 1 < w) one(src, x + 1, y, dst);
            return;
        }
        const float *row = src + (long long)(y - R) * w + (x - R);
        const float va = row[(long long)R * w + R], vb = row[(long long)R * w + R + 1];
        float na = 0.f, da = 0.f, nb = 0.f, db = 0.f;
        const float *rp = row;
#pragma unroll
        for (int i = 0; i < D; ++i, rp += w) {
#pragma unroll
            for (int j = 0; j <= D; ++j) {
                const bool in_a = j < D && (i - R) * (i - R) + (j - R) * (j - R) <= R * R;
                const bool in_b = j >= 1 && (i - R) * (i - R) + (j - 1 - R) * (j - 1 - R) <= R * R;
                if (!in_a && !in_b) continue;
                const float v = rp[j];
                if (in_a) tap(v, va, i * D + j, na, da);
                if (in_b) tap(v, vb, i * D + j - 1, nb, db);
            }
        }
        dst[(long long)y * w + x] = ss_div(na, da);
        dst[(long long)y * w + x + 1] = ss_div(nb, db);
    }
};

// One pixel per call (the variant the pair version is measured against: SSDM_BILATERAL_PAIR=0).
template <int R>
struct DmBilateralSingle {
    enum { D = 2 * R + 1 };
    const float *in; int h, w; const unsigned *stats; float gmax; float *out;
    float s0[D * D], s1[D * D], s2[D * D];
    SS_MDEV void operator()(int x, int y, int z) const
    {
        const long long plane = (long long)h * w;
        const float *src = in + (long long)z * plane;
        const float v0 = src[(long long)y * w + x];
        float mn = __uint_as_float_compat(stats[2 * z]), mx = __uint_as_float_compat(stats[2 * z + 1]);
        if (gmax > 0.f) { mn = ss_div(mn, gmax); mx = ss_div(mx, gmax); }
        if (fabsf(mn - mx) < 1.1920929e-07f) { out[(long long)z * plane + (long long)y * w + x] = v0; return; }
        float num = 0.f, den = 0.f;
        if (x >= R && x < w - R && y >= R && y < h - R) {
            const float *row = src + (long long)(y - R) * w + (x - R);
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    if ((i - R) * (i - R) + (j - R) * (j - R) > R * R) continue;
                    const float v = row[(long long)i * w + j];
                    const float dv = ss_add(v, -v0);
                    const float q = ss_mul(dv, dv);
                    const float wgt = ss_fma(q, ss_fma(q, s2[i * D + j], s1[i * D + j]), s0[i * D + j]);
                    num = ss_fma(v, wgt, num);
                    den = ss_add(den, wgt);
                }
            }
        } else {
            for (int i = 0; i < D; ++i) {
                const float *row = src + (long long)dm_reflect101(y - R + i, h) * w;
                for (int j = 0; j < D; ++j) {
                    if ((i - R) * (i - R) + (j - R) * (j - R) > R * R) continue;
                    const float v = row[dm_reflect101(x - R + j, w)];
                    const float dv = ss_add(v, -v0);
                    const float q = ss_mul(dv, dv);
                    const float wgt = ss_fma(q, ss_fma(q, s2[i * D + j], s1[i * D + j]), s0[i * D + j]);
                    num = ss_fma(v, wgt, num);
                    den = ss_add(den, wgt);
                }
            }
        }
        out[(long long)z * plane + (long long)y * w + x] = ss_div(num, den);
    }
};

// get_focus_map (depth_map.py:53-62), in place over the n planes of e; frames summed in stack order.
struct DmFocusMap {
    float *e; int h, w; int n; int softmax; float temperature;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const long long plane = (long long)h * w;
        float *p = e + (long long)y * w + x;
        if (!softmax) {
            float s = p[0];
            for (int i = 1; i < n; ++i) s = ss_add(s, p[(long long)i * plane]);
            for (int i = 0; i < n; ++i) {
                float *q = p + (long long)i * plane;
                *q = s != 0.f ? ss_div(*q, s) : 0.f;
            }
        } else {
            float m = p[0];
            for (int i = 1; i < n; ++i) m = fmaxf(m, p[(long long)i * plane]);
            float s = 0.f;
            for (int i = 0; i < n; ++i) {
                float *q = p + (long long)i * plane;
                const float rel = expf(ss_div(ss_add(*q, -m), temperature));
                *q = rel;
                s = i ? ss_add(s, rel) : rel;
            }
            for (int i = 0; i < n; ++i) {
                float *q = p + (long long)i * plane;
                *q = ss_div(*q, s);
            }
        }
    }
};

// ---- pass 2, per frame: Gaussian pyramids of the image and of its weight map, Laplacian blend -----------
// Source of one pyramid level: planes 0..2 = B, G, R (level 0: the interleaved integer frame, converted
// on the fly = img.astype(float32), depth_map.py:96; deeper levels: float planes), plane 3 = weight map.
template <typename InT>
struct DmLevelSrc {
    const void *frame; long long pitch;       // level 0 image (InT != float)
    const float *planes; const float *weight; // float planes (3, contiguous) / weight plane
    int h, w;
    SS_MDEV float at(int c, int x, int y) const
    {
        if (c == 3) return weight[(long long)y * w + x];
        if (frame) return (float)((const InT *)((const char *)frame + (long long)y * pitch))[3 * x + c];
        return planes[((long long)c * h + y) * w + x];
    }
};

// cv2.pyrDown (depth_map.py:101-102): row = 6*s[2x] + 4*(s[2x-1] + s[2x+1]) + s[2x-2] + s[2x+2] on five
// source rows (REFLECT101), then the same vertically, * 1/256.  One call computes all four planes (B, G, R,
// weight) of one output pixel: the ten reflected source indices are shared by the planes.
template <typename P>
SS_DEV float dm_pyrdown_px(const P *const *rows, const int *xo)
{
    float r[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const P *p = rows[i];
        const float s1 = ss_add((float)p[xo[1]], (float)p[xo[2]]);
        float acc = ss_add(ss_mul((float)p[xo[0]], 6.f), ss_mul(s1, 4.f));
        acc = ss_add(acc, (float)p[xo[3]]);
        r[i] = ss_add(acc, (float)p[xo[4]]);
    }
    float acc = ss_add(ss_mul(r[0], 6.f), ss_mul(ss_add(r[1], r[2]), 4.f));
    acc = ss_add(acc, r[3]);
    acc = ss_add(acc, r[4]);
    return ss_mul(acc, 0.00390625f);
}

template <typename InT>
struct DmPyrDown {
    DmLevelSrc<InT> src; int h2, w2; float *out /* 4 planes h2 x w2 */;
    SS_MDEV void operator()(int x2, int y2, int) const
    {
        const int h = src.h, w = src.w, x = 2 * x2, y = 2 * y2;
        int xs[5], ys[5];   // order of use: centre, -1, +1, -2, +2
        if (x >= 2 && x + 2 < w) { xs[0] = x; xs[1] = x - 1; xs[2] = x + 1; xs[3] = x - 2; xs[4] = x + 2; }
        else { xs[0] = x; xs[1] = dm_reflect101(x - 1, w); xs[2] = dm_reflect101(x + 1, w);
               xs[3] = dm_reflect101(x - 2, w); xs[4] = dm_reflect101(x + 2, w); }
        if (y >= 2 && y + 2 < h) { ys[0] = y; ys[1] = y - 1; ys[2] = y + 1; ys[3] = y - 2; ys[4] = y + 2; }
        else { ys[0] = y; ys[1] = dm_reflect101(y - 1, h); ys[2] = dm_reflect101(y + 1, h);
               ys[3] = dm_reflect101(y - 2, h); ys[4] = dm_reflect101(y + 2, h); }
        const long long o = (long long)y2 * w2 + x2, plane2 = (long long)h2 * w2;
        if (src.frame) {    // level 0: interleaved integer frame, the three channels share row pointers
            const InT *rows[5];
            int xo[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                rows[i] = (const InT *)((const char *)src.frame + (long long)ys[i] * src.pitch);
                xo[i] = 3 * xs[i];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                out[c * plane2 + o] = dm_pyrdown_px<InT>(rows, xo);
#pragma unroll
                for (int i = 0; i < 5; ++i) ++rows[i];
            }
        } else {
            const long long plane = (long long)h * w;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float *rows[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) rows[i] = src.planes + c * plane + (long long)ys[i] * w;
                out[c * plane2 + o] = dm_pyrdown_px<float>(rows, xs);
            }
        }
        const float *rows[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) rows[i] = src.weight + (long long)ys[i] * w;
        out[3 * plane2 + o] = dm_pyrdown_px<float>(rows, xs);
    }
};

// cv2.pyrUp(src, dstsize) (depth_map.py:106, :120) at destination pixel (x, y) of a coarse float plane
// h2 x w2: even columns l + 6c + r, odd columns 4(c + r); coarse index -1 -> 1, n -> n - 1; then the same
// vertically, * 1/64.  DmUpIdx holds the coarse indices of one destination pixel (shared by the channels).
struct DmUpIdx {
    int xl, xc, xr, odd_x; long long yu, yc, yd; int odd_y;
};
SS_DEV DmUpIdx dm_pyrup_idx(int h2, int w2, int x, int y)
{
    DmUpIdx i;
    i.xc = x >> 1; i.odd_x = x & 1;
    i.xr = i.xc + 1 < w2 ? i.xc + 1 : w2 - 1;
    i.xl = i.xc > 0 ? i.xc - 1 : (w2 > 1 ? 1 : 0);
    const int yc = y >> 1;
    i.odd_y = y & 1;
    i.yc = (long long)yc * w2;
    i.yd = (long long)(yc + 1 < h2 ? yc + 1 : h2 - 1) * w2;
    i.yu = (long long)(yc > 0 ? yc - 1 : (h2 > 1 ? 1 : 0)) * w2;
    return i;
}
SS_DEV float dm_pyrup_row(const float *line, const DmUpIdx &i)
{
    const float c = line[i.xc], r = line[i.xr];
    if (i.odd_x) return ss_mul(ss_add(c, r), 4.f);
    return ss_add(ss_add(line[i.xl], ss_mul(c, 6.f)), r);
}
SS_DEV float dm_pyrup_at(const float *coarse, const DmUpIdx &i)
{
    const float c = dm_pyrup_row(coarse + i.yc, i), d = dm_pyrup_row(coarse + i.yd, i);
    if (i.odd_y) return ss_mul(ss_mul(ss_add(c, d), 4.f), 0.015625f);
    const float u = dm_pyrup_row(coarse + i.yu, i);
    return ss_mul(ss_add(ss_add(u, ss_mul(c, 6.f)), d), 0.015625f);
}

// One level of the frame's weighted Laplacian pyramid added to the running blend (depth_map.py:104-111):
// acc[c] (+)= (g_l[c] - pyrUp(g_{l+1}[c])) * w_l, or g_l[c] * w_l at the coarsest level (coarse == NULL).
// The first frame stores instead of adding.  One call handles the three channels of one pixel.
template <typename InT>
struct DmBlendLevel {
    DmLevelSrc<InT> fine; const float *coarse /* 4 planes h2 x w2 or NULL */; int h2, w2;
    float *acc /* 3 planes */; int first;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const float wgt = fine.at(3, x, y);
        DmUpIdx ix;
        if (coarse) ix = dm_pyrup_idx(h2, w2, x, y);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = fine.at(c, x, y);
            if (coarse) v = ss_add(v, -dm_pyrup_at(coarse + (long long)c * h2 * w2, ix));
            v = ss_mul(v, wgt);
            float *a = acc + ((long long)c * fine.h + y) * fine.w + x;
            *a = first ? v : ss_add(*a, v);
        }
    }
};

// One collapse step (depth_map.py:118-120): fine[c] = pyrUp(coarse[c]) + fine[c], in place.
struct DmCollapse {
    const float *coarse; int h2, w2; float *fine; int h, w;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const DmUpIdx ix = dm_pyrup_idx(h2, w2, x, y);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float *a = fine + ((long long)c * h + y) * w + x;
            *a = ss_add(dm_pyrup_at(coarse + (long long)c * h2 * w2, ix), *a);
        }
    }
};

// np.clip(np.absolute(result), 0, n_values).astype(dtype) (depth_map.py:121-123) into interleaved BGR.
template <typename OutT>
struct DmStoreOutput {
    const float *res /* 3 planes */; int h, w; float maxv; void *out; long long pitch;
    SS_MDEV void operator()(int x, int y, int) const
    {
        OutT *o = (OutT *)((char *)out + (long long)y * pitch) + 3 * x;
        for (int c = 0; c < 3; ++c) {
            float v = fabsf(res[((long long)c * h + y) * w + x]);
            v = v > maxv ? maxv : v;
            o[c] = (OutT)(int)v;
        }
    }
};

#ifndef SS_HOST_EMU
template <class K>
__global__ void __launch_bounds__(256) dm_px_kernel(const __grid_constant__ K k, int w, int h)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < w && y < h) k(x, y, (int)blockIdx.z);
}
#endif

#ifndef SS_HOST_EMU
// persistent grid, one warp per list entry; the entry count is read from device memory (no host round trip)
template <class K>
__global__ void __launch_bounds__(256) dm_item_kernel(const __grid_constant__ K k, const int *count)
{
    const int n = *count, lane = threadIdx.x & 31;
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) k.item(i, lane, 32);
}
#endif
