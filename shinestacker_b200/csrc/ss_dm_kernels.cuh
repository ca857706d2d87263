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

// ---- pass 1 as ONE tiled kernel ----------------------------------------------------------------------------
// Kernels of this kind are functors "one call = one tile": tile(tx, ty, z, smem, tid, nthreads) runs stages
// separated by DM_SYNC(); inside a stage the work is split by (tid, nthreads) only, so the product runs it with
// a CTA of 256 threads and the host build with tid = 0, nthreads = 1 (DM_SYNC() empty) -- same arithmetic.
#ifdef SS_HOST_EMU
#define DM_SYNC() ((void)0)
#else
#define DM_SYNC() __syncthreads()
#endif
#define DM_ET_X 64
#define DM_ET_Y 32

// four consecutive floats of shared memory (16-byte aligned) as one access
struct DmF4 { float v[4]; };
#ifdef SS_HOST_EMU
SS_DEV DmF4 dm_ld4(const float *p) { DmF4 r; memcpy(r.v, p, 16); return r; }
SS_DEV void dm_st4(float *p, const DmF4 &a) { memcpy(p, a.v, 16); }
SS_DEV uint32_t dm_ldg32(const void *p) { uint32_t v; memcpy(&v, p, 4); return v; }
#else
SS_DEV DmF4 dm_ld4(const float *p) { const float4 t = *(const float4 *)p; DmF4 r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r; }
SS_DEV void dm_st4(float *p, const DmF4 &a) { *(float4 *)p = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }
SS_DEV uint32_t dm_ldg32(const void *p) { return __ldg((const uint32_t *)p); }
#endif
// b or -b: sign flipped by mask 0x80000000 (anti-symmetric kernels pair x[+j] - x[-j])
SS_DEV float dm_flip(float b, unsigned mask) { return __uint_as_float_compat(__float_as_uint_compat(b) ^ mask); }

// symmetric-form tap sum around p[0] (neighbours at p[+-j*stride]); REFL: positions reflected into [0, n)
template <int R, bool REFL>
SS_DEV float dm_sm_filter(const float *p, int stride, int pos, int n, const DmTaps &t)
{
    float acc = ss_mul(p[0], t.k[0]);
    const int r = R >= 0 ? R : t.r;
    const unsigned flip = t.anti ? 0x80000000u : 0u;
#pragma unroll
    for (int j = 1; j <= r; ++j) {
        int dp = j, dn = -j;
        if (REFL) { dp = dm_reflect101(pos + j, n) - pos; dn = dm_reflect101(pos - j, n) - pos; }
        acc = ss_add(acc, ss_mul(ss_add(p[dp * stride], dm_flip(p[dn * stride], flip)), t.k[j]));
    }
    return acc;
}

// the same sum for four neighbouring outputs: along a row from the 12 samples m[0..11] around them (outputs at
// m[4..7]), along a column from the rows at p[+-j*stride]
template <int R>
SS_DEV DmF4 dm_row4(const float *m, const DmTaps &t)
{
    const unsigned flip = t.anti ? 0x80000000u : 0u;
    DmF4 r;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float acc = ss_mul(m[4 + o], t.k[0]);
#pragma unroll
        for (int j = 1; j <= R; ++j) acc = ss_add(acc, ss_mul(ss_add(m[4 + o + j], dm_flip(m[4 + o - j], flip)), t.k[j]));
        r.v[o] = acc;
    }
    return r;
}
template <int R>
SS_DEV DmF4 dm_col4(const float *p, int stride, const DmTaps &t)
{
    const unsigned flip = t.anti ? 0x80000000u : 0u;
    DmF4 r = dm_ld4(p);
#pragma unroll
    for (int o = 0; o < 4; ++o) r.v[o] = ss_mul(r.v[o], t.k[0]);
#pragma unroll
    for (int j = 1; j <= R; ++j) {
        const DmF4 a = dm_ld4(p + j * stride), b = dm_ld4(p - j * stride);
#pragma unroll
        for (int o = 0; o < 4; ++o) r.v[o] = ss_add(r.v[o], ss_mul(ss_add(a.v[o], dm_flip(b.v[o], flip)), t.k[j]));
    }
    return r;
}

// The four passes above (DmGrayRowFilter -> DmColFilter -> DmRowPair -> DmColCombine) for one 64 x 32 tile of
// one frame with every intermediate in shared memory: the frame is read once, the energy written once.  All
// stages work in IMAGE coordinates on windows clipped to the image and reflect their taps (REFLECT101 of the
// stage's own input plane, as cv2 applies it to each filter call), so border tiles need no special casing;
// the tap order of every filter is that of dm_filter1d, hence the result is bit-identical to the four passes.
// Tiles whose window lies inside the image take `fast`: four pixels per work item, 128-bit shared accesses
// (slots outside a stage's valid window are computed from whatever lies there and never consumed).
// RB / RD: radii of the blur and derivative kernels when known at compile time (-1: read from the taps).
// Patch: ph = 32 + 2 (rb + rd) rows of pw = 64 + 2 hx floats, hx = rb + rd rounded up to 4; three planes behind a
// 4-float pad each (the 128-bit row accesses of `fast` reach one group past either end of a plane; those slots are
// never written and what is read there is never used): DM_ET_FLOATS(pw, ph).
#define DM_ET_FLOATS(pw, ph) (3 * ((pw) * (ph) + 4) + 4)
template <typename InT, int RB, int RD>
struct DmEnergyTile {
    const void *frame; long long pitch; int h, w; DmTaps tblur, ta, tb; int sobel; float *out; unsigned *stats;
    int pw, ph;

    SS_MDEV float combine(float da, float db) const { return sobel ? ss_add(fabsf(da), fabsf(db)) : fabsf(ss_add(da, db)); }

    template <bool REFL>
    SS_MDEV void run(int x0, int y0, float *sm, int tid, int nthreads, unsigned &lo, unsigned &hi) const
    {
        const int rb = RB >= 0 ? RB : tblur.r, rd = RD >= 0 ? RD : (ta.r > tb.r ? ta.r : tb.r);
        const int H = rb + rd, ox = x0 - (pw - DM_ET_X) / 2, oy = y0 - H;
        const int lanes = nthreads < 32 ? nthreads : 32, lane = tid % lanes, warp = tid / lanes, nwarps = nthreads / lanes;
        const int gx0 = x0 - H > 0 ? x0 - H : 0, gx1 = x0 + DM_ET_X + H < w ? x0 + DM_ET_X + H : w;
        const int gy0 = oy > 0 ? oy : 0, gy1 = y0 + DM_ET_Y + H < h ? y0 + DM_ET_Y + H : h;
        const int bx0 = x0 - rd > 0 ? x0 - rd : 0, bx1 = x0 + DM_ET_X + rd < w ? x0 + DM_ET_X + rd : w;
        const int by0 = y0 - rd > 0 ? y0 - rd : 0, by1 = y0 + DM_ET_Y + rd < h ? y0 + DM_ET_Y + rd : h;
        const int ex1 = x0 + DM_ET_X < w ? x0 + DM_ET_X : w, ey1 = y0 + DM_ET_Y < h ? y0 + DM_ET_Y : h;
        float *P0 = sm + 4, *P1 = P0 + pw * ph + 4, *P2 = P1 + pw * ph + 4;
#define DM_AT(P, y, x) ((P) + ((y) - oy) * pw + ((x) - ox))
        // G: gray of the window
        for (int y = gy0 + warp; y < gy1; y += nwarps)
            for (int x = gx0 + lane; x < gx1; x += lanes) *DM_AT(P0, y, x) = dm_gray_at<InT>(frame, pitch, x, y);
        DM_SYNC();
        if (rb > 0) {
            // R: rows of the blur, P0 -> P1
            for (int y = gy0 + warp; y < gy1; y += nwarps)
                for (int x = bx0 + lane; x < bx1; x += lanes)
                    *DM_AT(P1, y, x) = dm_sm_filter<RB, REFL>(DM_AT(P0, y, x), 1, x, w, tblur);
            DM_SYNC();
            // B: columns of the blur, P1 -> P0
            for (int y = by0 + warp; y < by1; y += nwarps)
                for (int x = bx0 + lane; x < bx1; x += lanes)
                    *DM_AT(P0, y, x) = dm_sm_filter<RB, REFL>(DM_AT(P1, y, x), pw, y, h, tblur);
            DM_SYNC();
        }
        // D: rows of the derivative pair, P0 -> P1 (ta), P2 (tb)
        for (int y = by0 + warp; y < by1; y += nwarps)
            for (int x = x0 + lane; x < ex1; x += lanes) {
                *DM_AT(P1, y, x) = dm_sm_filter<RD, REFL>(DM_AT(P0, y, x), 1, x, w, ta);
                *DM_AT(P2, y, x) = dm_sm_filter<RD, REFL>(DM_AT(P0, y, x), 1, x, w, tb);
            }
        DM_SYNC();
        // E: columns + combination
        for (int y = y0 + warp; y < ey1; y += nwarps)
            for (int x = x0 + lane; x < ex1; x += lanes) {
                const float da = dm_sm_filter<RD, REFL>(DM_AT(P1, y, x), pw, y, h, tb);
                const float db = dm_sm_filter<RD, REFL>(DM_AT(P2, y, x), pw, y, h, ta);
                const float e = combine(da, db);
                out[(long long)y * w + x] = e;
                const unsigned bits = __float_as_uint_compat(e);
                lo = bits < lo ? bits : lo; hi = bits > hi ? bits : hi;
            }
#undef DM_AT
    }

    // interior tile, compile-time radii, image width a multiple of 4
    SS_MDEV void fast(int x0, int y0, float *sm, int tid, int nthreads, unsigned &lo, unsigned &hi) const
    {
        enum { RBc = RB > 0 ? RB : 0, RDc = RD > 0 ? RD : 0, H = RBc + RDc, HX = (H + 3) / 4 * 4, PW = DM_ET_X + 2 * HX,
               PH = DM_ET_Y + 2 * H, G = PW / 4 };
        float *P0 = sm + 4, *P1 = P0 + PW * PH + 4, *P2 = P1 + PW * PH + 4;
        const int ox = x0 - HX, oy = y0 - H;
        // G: gray, four pixels (12 samples) per item
        for (int item = tid; item < PH * G; item += nthreads) {
            const int i = item / G, g = item % G;
            const char *src = (const char *)frame + (long long)(oy + i) * pitch + (long long)(ox + 4 * g) * 3 * (int)sizeof(InT);
            uint32_t wd[3 * sizeof(InT)];
#pragma unroll
            for (int k = 0; k < 3 * (int)sizeof(InT); ++k) wd[k] = dm_ldg32(src + 4 * k);
            DmF4 r;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                uint32_t c[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int q = 3 * o + k;
                    c[k] = sizeof(InT) == 1 ? (wd[q >> 2] >> (8 * (q & 3))) & 0xffu : (wd[q >> 1] >> (16 * (q & 1))) & 0xffffu;
                }
                r.v[o] = (float)((c[0] * 3735u + c[1] * 19235u + c[2] * 9798u + 16384u) >> 15);
            }
            dm_st4(P0 + i * PW + 4 * g, r);
        }
        DM_SYNC();
        if (RBc > 0) {
            for (int item = tid; item < PH * G; item += nthreads) {       // R: P0 -> P1
                const int i = item / G, g = item % G;
                float m[12];
                const float *p = P0 + i * PW + 4 * g - 4;
#pragma unroll
                for (int q = 0; q < 3; ++q) { const DmF4 t = dm_ld4(p + 4 * q); m[4 * q] = t.v[0]; m[4 * q + 1] = t.v[1]; m[4 * q + 2] = t.v[2]; m[4 * q + 3] = t.v[3]; }
                dm_st4(P1 + i * PW + 4 * g, dm_row4<RBc>(m, tblur));
            }
            DM_SYNC();
            for (int item = tid; item < (PH - 2 * RBc) * G; item += nthreads) {    // B: P1 -> P0
                const int i = RBc + item / G, g = item % G;
                dm_st4(P0 + i * PW + 4 * g, dm_col4<RBc>(P1 + i * PW + 4 * g, PW, tblur));
            }
            DM_SYNC();
        }
        for (int item = tid; item < (PH - 2 * RBc) * (DM_ET_X / 4); item += nthreads) {       // D: P0 -> P1, P2
            const int i = RBc + item / (DM_ET_X / 4), g = HX / 4 + item % (DM_ET_X / 4);
            float m[12];
            const float *p = P0 + i * PW + 4 * g - 4;
#pragma unroll
            for (int q = 0; q < 3; ++q) { const DmF4 t = dm_ld4(p + 4 * q); m[4 * q] = t.v[0]; m[4 * q + 1] = t.v[1]; m[4 * q + 2] = t.v[2]; m[4 * q + 3] = t.v[3]; }
            dm_st4(P1 + i * PW + 4 * g, dm_row4<RDc>(m, ta));
            dm_st4(P2 + i * PW + 4 * g, dm_row4<RDc>(m, tb));
        }
        DM_SYNC();
        for (int item = tid; item < DM_ET_Y * (DM_ET_X / 4); item += nthreads) {               // E
            const int i = H + item / (DM_ET_X / 4), g = HX / 4 + item % (DM_ET_X / 4);
            const DmF4 da = dm_col4<RDc>(P1 + i * PW + 4 * g, PW, tb), db = dm_col4<RDc>(P2 + i * PW + 4 * g, PW, ta);
            DmF4 e;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                e.v[o] = combine(da.v[o], db.v[o]);
                const unsigned bits = __float_as_uint_compat(e.v[o]);
                lo = bits < lo ? bits : lo; hi = bits > hi ? bits : hi;
            }
            dm_st4(out + (long long)(oy + i) * w + ox + 4 * g, e);
        }
    }

    SS_MDEV void tile(int tx, int ty, int, float *sm, int tid, int nthreads) const
    {
        const int rb = RB >= 0 ? RB : tblur.r, rd = RD >= 0 ? RD : (ta.r > tb.r ? ta.r : tb.r);
        const int H = rb + rd, HX = (pw - DM_ET_X) / 2, x0 = tx * DM_ET_X, y0 = ty * DM_ET_Y;
        unsigned lo = 0x7f800000u, hi = 0u;
        const bool inside = x0 - HX >= 0 && y0 - H >= 0 && x0 + DM_ET_X + HX <= w && y0 + DM_ET_Y + H <= h;
        if (RB >= 0 && RD >= 0 && inside && (w & 3) == 0) fast(x0, y0, sm, tid, nthreads, lo, hi);
        else if (inside) run<false>(x0, y0, sm, tid, nthreads, lo, hi);
        else run<true>(x0, y0, sm, tid, nthreads, lo, hi);
        // per-frame minimum / maximum: one pair of atomics per warp and tile (see dm_stats_update)
#ifdef SS_HOST_EMU
        if (lo < stats[0]) stats[0] = lo;
        if (hi > stats[1]) stats[1] = hi;
#else
        unsigned *slot = stats + 2 * ((blockIdx.x + 7 * blockIdx.y) & (DM_STAT_SLOTS - 1));
        lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
        if ((tid & 31) == 0) {
            if (lo < __ldcg(slot)) atomicMin(slot, lo);
            if (hi > __ldcg(slot + 1)) atomicMax(slot + 1, hi);
        }
#endif
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

// ---- the default bilateral (radius 7) as a tiled kernel in moment form ------------------------------------
// With the colour weight 1 + c1 d^2 + c2 d^4, d = v_q - v_p (see DmBilateral), numerator and denominator are
// polynomials in v_p whose coefficients are plain convolutions of the powers of the plane with the space kernel:
//     M_k(p) = sum_q s_q v_q^k,     A_0..A_4 = 1 + c1 v^2 + c2 v^4, -2 c1 v - 4 c2 v^3, c1 + 6 c2 v^2, -4 c2 v, c2
//     den = sum_j A_j M_j,          num = sum_j A_j M_{j+1}                                  (v = v_p)
// and the space kernel, a Gaussian cut to a disc, is g(dx) g(dy) on rows of half-width 7,6,6,6,5,4,3,0 for
// |dy| = 0..7: a row pass forms the nested partial sums H_0, H_3..H_7 of every patch row (15 taps for all six),
// a column pass picks H_{hw(dy)} at row y + dy (15 taps) -- 30 FMA per moment and pixel instead of 149 x 6
// operations per pixel.  Cancellation is harmless: the brackets that cancel are scaled by c1 = -8e-4 and
// c2 = 3.2e-7, so their float32 rounding stays below 1e-10 of the result; the space weight g(dx) g(dy) differs
// from cv2's table entry by one rounding.  Contract: focus maps within 5e-6 of the oracle (tests), like
// DmBilateralFixed.  Borders: REFLECT101 of the plane (cv2.bilateralFilter's BORDER_DEFAULT) when the patch loads.
#define DM_BT_X 64
#define DM_BT_Y 32
#define DM_BT_PP 80                        /* patch pitch: 64 + 14 -> 80 floats (16-byte row groups) */
#define DM_BT_PH (DM_BT_Y + 14)
#define DM_BT_FLOATS (DM_BT_PH * DM_BT_PP + 6 * DM_BT_PH * DM_BT_X)
#ifdef SS_HOST_EMU
#define DM_BT_ITEMS 256                    /* work items one (emulated) thread walks through */
#else
#define DM_BT_ITEMS 1
#endif

template <int K>
SS_DEV float dm_pow(float v)
{
    if (K == 1) return v;
    const float v2 = ss_mul(v, v);
    if (K == 2) return v2;
    if (K == 3) return ss_mul(v2, v);
    const float v4 = ss_mul(v2, v2);
    if (K == 4) return v4;
    return ss_mul(v4, v);
}

// KMAX = 5: the full polynomial; KMAX = 3: c2 = 0, i.e. exp(t) = 1 + t, which is off by at most t^2/2 <= 3.2e-7
// relative in the weights -- below the float32 rounding of either form (measured: both within 5e-7 of the oracle's
// focus maps) -- and needs only M_1..M_3.  norm > 0: the plane holds raw energies, divided by norm (the global
// maximum, DmNormalize) as the patch loads.
template <int KMAX>
struct DmBilateralTile {
    const float *in; int h, w; const unsigned *stats; float gmax; float *out;
    float g[8];             // exp(-d^2 / (2 * 25^2)), d = 0..7
    float c1, c2, m0;       // colour polynomial; m0 = sum of the space weights over the disc
    float norm;

    // row pass of moment K: items = patch rows x groups of four output columns
    template <int K>
    SS_MDEV void rows(const float *V, float *H, int tid, int nthreads) const
    {
        for (int item = tid; item < DM_BT_PH * (DM_BT_X / 4); item += nthreads) {
            const int i = item / (DM_BT_X / 4), x4 = 4 * (item % (DM_BT_X / 4));
            const float *src = V + i * DM_BT_PP + x4;
            float m[20];
#pragma unroll
            for (int q = 0; q < 5; ++q) {                                   // 128-bit loads (x4 and the pitch are multiples of 4)
                const DmF4 t = dm_ld4(src + 4 * q);
#pragma unroll
                for (int j = 0; j < 4; ++j) m[4 * q + j] = dm_pow<K>(t.v[j]);
            }
            float *dst = H + i * DM_BT_X + x4;
            DmF4 hv[6];                                                     // planes H_0, H_3, H_4, H_5, H_6, H_7
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int c = o + 7;
                hv[0].v[o] = m[c];
                float a = m[c];
#pragma unroll
                for (int d = 1; d <= 7; ++d) {
                    a = ss_fma(g[d], ss_add(m[c - d], m[c + d]), a);
                    if (d >= 3) hv[d - 2].v[o] = a;
                }
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) dm_st4(dst + q * DM_BT_PH * DM_BT_X, hv[q]);    // 128-bit stores
        }
    }

    // column pass of moment K: items = output columns x strips of eight rows; den += A_K M_K, num += A_{K-1} M_K
    template <int K>
    SS_MDEV void cols(const float *V, const float *H, float (*acc)[16], int tid, int nthreads) const
    {
        const int P = DM_BT_PH * DM_BT_X;
        for (int it = 0, item = tid; it < DM_BT_ITEMS && item < DM_BT_X * (DM_BT_Y / 8); ++it, item += nthreads) {
            const int x = item % DM_BT_X, r0 = 8 * (item / DM_BT_X) + 7;
            const float *hp = H + x;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int r = r0 + o;
                float mk = hp[5 * P + r * DM_BT_X];
#pragma unroll
                for (int d = 1; d <= 3; ++d)
                    mk = ss_fma(g[d], ss_add(hp[4 * P + (r - d) * DM_BT_X], hp[4 * P + (r + d) * DM_BT_X]), mk);
                mk = ss_fma(g[4], ss_add(hp[3 * P + (r - 4) * DM_BT_X], hp[3 * P + (r + 4) * DM_BT_X]), mk);
                mk = ss_fma(g[5], ss_add(hp[2 * P + (r - 5) * DM_BT_X], hp[2 * P + (r + 5) * DM_BT_X]), mk);
                mk = ss_fma(g[6], ss_add(hp[1 * P + (r - 6) * DM_BT_X], hp[1 * P + (r + 6) * DM_BT_X]), mk);
                mk = ss_fma(g[7], ss_add(hp[(r - 7) * DM_BT_X], hp[(r + 7) * DM_BT_X]), mk);
                const float v = V[r * DM_BT_PP + x + 7], v2 = ss_mul(v, v);
                const float a0 = ss_fma(v2, ss_fma(v2, c2, c1), 1.f);
                const float a1 = ss_mul(v, ss_fma(v2, -4.f * c2, -2.f * c1));
                const float a2 = ss_fma(v2, 6.f * c2, c1);
                const float a3 = ss_mul(v, -4.f * c2);
                const float ak = K == 1 ? a1 : K == 2 ? a2 : K == 3 ? a3 : c2;       // A_K   (unused for K = 5)
                const float an = K == 1 ? a0 : K == 2 ? a1 : K == 3 ? a2 : K == 4 ? a3 : c2;   // A_{K-1}
                if (K == 1) { acc[it][o] = ss_fma(ak, mk, ss_mul(a0, m0)); acc[it][8 + o] = ss_mul(an, mk); }
                else {
                    if (K < 5) acc[it][o] = ss_fma(ak, mk, acc[it][o]);
                    acc[it][8 + o] = ss_fma(an, mk, acc[it][8 + o]);
                }
            }
        }
    }

    SS_MDEV void tile(int tx, int ty, int z, float *sm, int tid, int nthreads) const
    {
        const long long plane = (long long)h * w;
        const float *src = in + (long long)z * plane;
        float *dst = out + (long long)z * plane;
        const int x0 = tx * DM_BT_X, y0 = ty * DM_BT_Y;
        const int lanes = nthreads < 32 ? nthreads : 32, lane = tid % lanes, warp = tid / lanes, nwarps = nthreads / lanes;
        float mn = __uint_as_float_compat(stats[2 * z]), mx = __uint_as_float_compat(stats[2 * z + 1]);
        if (gmax > 0.f) { mn = ss_div(mn, gmax); mx = ss_div(mx, gmax); }
        if (fabsf(mn - mx) < 1.1920929e-07f) {      // cv2 copies a flat plane
            for (int y = y0 + warp; y < y0 + DM_BT_Y && y < h; y += nwarps)
                for (int x = x0 + lane; x < x0 + DM_BT_X && x < w; x += lanes) {
                    const float v = src[(long long)y * w + x];
                    dst[(long long)y * w + x] = norm > 0.f ? ss_div(v, norm) : v;
                }
            return;
        }
        float *V = sm, *H = sm + DM_BT_PH * DM_BT_PP;
        const bool inside = x0 >= 7 && y0 >= 7 && x0 + DM_BT_X + 7 <= w && y0 + DM_BT_Y + 7 <= h;
        // patch: DM_BT_PH x DM_BT_PP slots, a batch of loads in flight per thread before the stores
        for (int base = 0; base < DM_BT_PH * DM_BT_PP; base += 8 * nthreads) {
            float t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int s = base + u * nthreads + tid, i = s / DM_BT_PP, j = s % DM_BT_PP;
                t[u] = 0.f;
                if (s < DM_BT_PH * DM_BT_PP && j < DM_BT_X + 14) {
                    const int yy = inside ? y0 - 7 + i : dm_reflect101(y0 - 7 + i, h);
                    const int xx = inside ? x0 - 7 + j : dm_reflect101(x0 - 7 + j, w);
                    t[u] = src[(long long)yy * w + xx];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int s = base + u * nthreads + tid;
                if (s < DM_BT_PH * DM_BT_PP) V[s] = norm > 0.f ? ss_div(t[u], norm) : t[u];     // after the loads: keeps them in flight
            }
        }
        DM_SYNC();
        float acc[DM_BT_ITEMS][16];
        rows<1>(V, H, tid, nthreads); DM_SYNC(); cols<1>(V, H, acc, tid, nthreads); DM_SYNC();
        rows<2>(V, H, tid, nthreads); DM_SYNC(); cols<2>(V, H, acc, tid, nthreads); DM_SYNC();
        rows<3>(V, H, tid, nthreads); DM_SYNC(); cols<3>(V, H, acc, tid, nthreads);
        if (KMAX >= 5) {
            DM_SYNC();
            rows<4>(V, H, tid, nthreads); DM_SYNC(); cols<4>(V, H, acc, tid, nthreads); DM_SYNC();
            rows<5>(V, H, tid, nthreads); DM_SYNC(); cols<5>(V, H, acc, tid, nthreads);
        }
        for (int it = 0, item = tid; it < DM_BT_ITEMS && item < DM_BT_X * (DM_BT_Y / 8); ++it, item += nthreads) {
            const int x = x0 + item % DM_BT_X, yb = y0 + 8 * (item / DM_BT_X);
            if (x >= w) continue;
#pragma unroll
            for (int o = 0; o < 8; ++o)
                if (yb + o < h) dst[(long long)(yb + o) * w + x] = ss_div(acc[it][8 + o], acc[it][o]);
        }
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

// ---- frames sharded over ranks (one process per GPU): the sums over the stack are taken rank by rank ----------
#define DM_MAX_RANKS 16
// sum of this rank's n planes, frames in stack order (get_focus_map 'average', depth_map.py:55-57)
struct DmPartialSum {
    const float *e; int h, w, n; float *out;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const long long plane = (long long)h * w, o = (long long)y * w + x;
        float s = e[o];
        for (int i = 1; i < n; ++i) s = ss_add(s, e[i * plane + o]);
        out[o] = s;
    }
};
// 'max' map (softmax, depth_map.py:58-60): per-pixel maximum of this rank's planes ...
struct DmPartialMax {
    const float *e; int h, w, n; float *out;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const long long plane = (long long)h * w, o = (long long)y * w + x;
        float m = e[o];
        for (int i = 1; i < n; ++i) m = fmaxf(m, e[i * plane + o]);
        out[o] = m;
    }
};
// ... then, with the maximum over every rank's planes (exact, whatever the order): the own planes become
// exp((e - max) / temperature) in place and their sum in stack order goes to the partial-sum plane
struct DmSoftmaxRel {
    float *e; int h, w, n; const float *pmax[DM_MAX_RANKS]; int world; float temperature; float *psum;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const long long plane = (long long)h * w, o = (long long)y * w + x;
        float m = pmax[0][o];
        for (int r = 1; r < world; ++r) m = fmaxf(m, pmax[r][o]);
        float s = 0.f;
        for (int i = 0; i < n; ++i) {
            float *q = e + i * plane + o;
            const float rel = expf(ss_div(ss_add(*q, -m), temperature));
            *q = rel;
            s = i ? ss_add(s, rel) : rel;
        }
        psum[o] = s;
    }
};
// sum of every rank's partial sum, ranks in stack order, then e_i / sum for the own n planes, in place
// (softmax: the division is unconditional, as in DmFocusMap)
struct DmFocusMapShared {
    float *e; int h, w, n; const float *psum[DM_MAX_RANKS]; int world; int softmax;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const long long plane = (long long)h * w, o = (long long)y * w + x;
        float s = psum[0][o];
        for (int r = 1; r < world; ++r) s = ss_add(s, psum[r][o]);
        for (int i = 0; i < n; ++i) {
            float *q = e + i * plane + o;
            *q = softmax ? ss_div(*q, s) : (s != 0.f ? ss_div(*q, s) : 0.f);
        }
    }
};
// One level of the blended pyramid, reduced in two steps.  Rows [rank h / world, (rank + 1) h / world) are this rank's
// band: DmPeerBandSum adds every rank's share of the band, ranks in stack order, into the own pyramid (in place: the
// other ranks read other rows of it); DmPeerBandGather then copies the finished bands of the other ranks into the
// root's pyramid.  One call = one float of one of the three planes.
struct DmPeerBandSum {
    const float *acc[DM_MAX_RANKS]; int world; float *out; int h, w, y0, y1;
    SS_MDEV void operator()(int x, int yy, int) const
    {
        const int rows = y1 - y0, plane = yy / rows, y = y0 + yy % rows;
        const long long o = ((long long)plane * h + y) * w + x;
        float s = acc[0][o];
        for (int r = 1; r < world; ++r) s = ss_add(s, acc[r][o]);
        out[o] = s;
    }
};
struct DmPeerBandGather {
    const float *acc[DM_MAX_RANKS]; int world, root; float *out; int h, w;
    SS_MDEV void operator()(int x, int yy, int) const
    {
        const int plane = yy / h, y = yy % h;
        int q = 0;
        while (q + 1 < world && y >= (int)((long long)(q + 1) * h / world)) ++q;
        if (q == root) return;
        const long long o = ((long long)plane * h + y) * w + x;
        out[o] = acc[q][o];
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
    long long zs_frame, zs_planes, zs_weight;   // distance between consecutive frames of a batch (bytes, floats, floats)
    SS_MDEV DmLevelSrc of(int z) const
    {
        DmLevelSrc s = *this;
        if (frame) s.frame = (const char *)frame + z * zs_frame;
        if (planes) s.planes = planes + z * zs_planes;
        s.weight = weight + z * zs_weight;
        return s;
    }
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
    DmLevelSrc<InT> src0; int h2, w2; float *out0 /* 4 planes h2 x w2 */; long long zs_out;
    SS_MDEV void operator()(int x2, int y2, int z) const       // z: frame of the batch
    {
        const DmLevelSrc<InT> src = src0.of(z);
        float *out = out0 + z * zs_out;
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

// cv2.pyrDown as a tiled kernel: one 32 x 16 tile of the coarse level for the four planes of one frame (z) of
// a batch.  The 67 x 35 source window (REFLECT101 per coordinate, like DmPyrDown) is loaded once into shared
// memory, rows then columns; every output is the same sequence of operations as dm_pyrdown_px.
#define DM_PD_X 32
#define DM_PD_Y 16
#define DM_PD_SW (2 * DM_PD_X + 4)
#define DM_PD_SH (2 * DM_PD_Y + 3)
#define DM_PD_FLOATS (4 * DM_PD_SH * DM_PD_SW + 4 * DM_PD_SH * DM_PD_X)
template <typename InT>
struct DmPyrDownTile {
    DmLevelSrc<InT> src0; int h2, w2; float *out0 /* 4 planes h2 x w2 */; long long zs_out;
    SS_MDEV void tile(int tx, int ty, int z, float *sm, int tid, int nthreads) const
    {
        const DmLevelSrc<InT> src = src0.of(z);
        float *out = out0 + z * zs_out;
        const int h = src.h, w = src.w, x0 = tx * DM_PD_X, y0 = ty * DM_PD_Y, sx0 = 2 * x0 - 2, sy0 = 2 * y0 - 2;
        float *S = sm, *R = sm + 4 * DM_PD_SH * DM_PD_SW;
        const bool inside = sx0 >= 0 && sy0 >= 0 && sx0 + DM_PD_SW - 1 <= w && sy0 + DM_PD_SH <= h;
        const int lanes = nthreads < 32 ? nthreads : 32, lane = tid % lanes, warp = tid / lanes, nwarps = nthreads / lanes;
        for (int i = warp; i < DM_PD_SH; i += nwarps) {         // 35 rows x 67 columns, a warp per row
            const int yy = inside ? sy0 + i : dm_reflect101(sy0 + i, h);
            float t[3][4];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int j = lane + u * lanes;
                if (j < DM_PD_SW - 1) {
                    const int xx = inside ? sx0 + j : dm_reflect101(sx0 + j, w);
#pragma unroll
                    for (int p = 0; p < 4; ++p) t[u][p] = src.at(p, xx, yy);
                }
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int j = lane + u * lanes;
                if (j < DM_PD_SW - 1) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) S[(p * DM_PD_SH + i) * DM_PD_SW + j] = t[u][p];
                }
            }
            for (int j = lane + 3 * lanes; j < DM_PD_SW - 1; j += lanes) {      // fewer than 32 lanes (host build)
                const int xx = inside ? sx0 + j : dm_reflect101(sx0 + j, w);
                for (int p = 0; p < 4; ++p) S[(p * DM_PD_SH + i) * DM_PD_SW + j] = src.at(p, xx, yy);
            }
        }
        DM_SYNC();
        // rows: two neighbouring outputs per item share their taps
        for (int item = tid; item < 4 * DM_PD_SH * (DM_PD_X / 2); item += nthreads) {
            const int x = 2 * (item % (DM_PD_X / 2)), pi = item / (DM_PD_X / 2);        // pi = plane * SH + row
            const float *s = S + pi * DM_PD_SW + 2 * x;
            float v[7];
#pragma unroll
            for (int q = 0; q < 7; ++q) v[q] = s[q];
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                float acc = ss_add(ss_mul(v[2 * o + 2], 6.f), ss_mul(ss_add(v[2 * o + 1], v[2 * o + 3]), 4.f));
                acc = ss_add(acc, v[2 * o]);
                R[pi * DM_PD_X + x + o] = ss_add(acc, v[2 * o + 4]);
            }
        }
        DM_SYNC();
        const long long plane2 = (long long)h2 * w2;
        for (int item = tid; item < 4 * DM_PD_Y * DM_PD_X; item += nthreads) {          // columns
            const int x = item % DM_PD_X, y = (item / DM_PD_X) % DM_PD_Y, p = item / (DM_PD_X * DM_PD_Y);
            if (x0 + x >= w2 || y0 + y >= h2) continue;
            const float *r = R + (p * DM_PD_SH + 2 * y + 2) * DM_PD_X + x;
            float acc = ss_add(ss_mul(r[0], 6.f), ss_mul(ss_add(r[-DM_PD_X], r[DM_PD_X]), 4.f));
            acc = ss_add(acc, r[-2 * DM_PD_X]);
            acc = ss_add(acc, r[2 * DM_PD_X]);
            out[p * plane2 + (long long)(y0 + y) * w2 + x0 + x] = ss_mul(acc, 0.00390625f);
        }
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
// The first frame stores instead of adding.  One call = the 2 x 2 block of pixels (2xi.., 2yi..), the three
// channels, the nf frames of a batch in stack order: the four pixels share one 3 x 3 neighbourhood of the coarse
// level (9 loads per channel instead of 25; same operations per pixel as dm_pyrup_at, so the same bits), and the
// running blend is read and written once per batch instead of once per frame.
#ifndef DM_BLEND_UNROLL
#define DM_BLEND_UNROLL 2       /* frames in flight per thread; A/B on B200 with 4 resident CTAs: 1 -> 48.3, 2 -> 43.0 ms/stack */
#endif
template <typename InT>
struct DmBlendLevel {
    DmLevelSrc<InT> fine0; const float *coarse0 /* 4 planes h2 x w2 or NULL */; int h2, w2; long long zs_coarse;
    float *acc /* 3 planes */; int first, nf;
    SS_MDEV void operator()(int xi, int yi, int) const
    {
        const int x = 2 * xi, y = 2 * yi, W = fine0.w, Hh = fine0.h;
        const bool ok[4] = {true, x + 1 < W, y + 1 < Hh, x + 1 < W && y + 1 < Hh};
        const long long plane = (long long)Hh * W;
        long long off[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) off[k] = (long long)(y + (k >> 1)) * W + x + (k & 1);
        float a[3][4];
        if (!first) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) a[c][k] = ok[k] ? acc[c * plane + off[k]] : 0.f;
        }
        DmUpIdx ix;
        if (coarse0) ix = dm_pyrup_idx(h2, w2, x, y);
        constexpr int unroll_frames = DM_BLEND_UNROLL;
#pragma unroll unroll_frames
        for (int f = 0; f < nf; ++f) {
            const DmLevelSrc<InT> fine = fine0.of(f);
            float wgt[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) wgt[k] = ok[k] ? fine.at(3, x + (k & 1), y + (k >> 1)) : 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = ok[k] ? fine.at(c, x + (k & 1), y + (k >> 1)) : 0.f;
                if (coarse0) {
                    const float *cp = coarse0 + f * zs_coarse + (long long)c * h2 * w2;
                    float ev[3], od[3];         // rows up / centre / down of the coarse level, even and odd column
                    const long long rows[3] = {ix.yu, ix.yc, ix.yd};
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float l = cp[rows[r] + ix.xl], m = cp[rows[r] + ix.xc], rr = cp[rows[r] + ix.xr];
                        ev[r] = ss_add(ss_add(l, ss_mul(m, 6.f)), rr);
                        od[r] = ss_mul(ss_add(m, rr), 4.f);
                    }
                    v[0] = ss_add(v[0], -ss_mul(ss_add(ss_add(ev[0], ss_mul(ev[1], 6.f)), ev[2]), 0.015625f));
                    v[1] = ss_add(v[1], -ss_mul(ss_add(ss_add(od[0], ss_mul(od[1], 6.f)), od[2]), 0.015625f));
                    v[2] = ss_add(v[2], -ss_mul(ss_mul(ss_add(ev[1], ev[2]), 4.f), 0.015625f));
                    v[3] = ss_add(v[3], -ss_mul(ss_mul(ss_add(od[1], od[2]), 4.f), 0.015625f));
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float t = ss_mul(v[k], wgt[k]);
                    a[c][k] = (first && f == 0) ? t : ss_add(a[c][k], t);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (ok[k]) acc[c * plane + off[k]] = a[c][k];
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
// resident CTAs per SM a pixel functor is compiled for (register cap = 65536 / (256 * v))
template <class K> struct DmMinBlocks { enum { v = 1 }; };
#ifndef DM_BLEND_MINBLOCKS
#define DM_BLEND_MINBLOCKS 4
#endif
template <typename InT> struct DmMinBlocks<DmBlendLevel<InT>> { enum { v = DM_BLEND_MINBLOCKS }; };
template <class K>
__global__ void __launch_bounds__(256, DmMinBlocks<K>::v) dm_px_kernel(const __grid_constant__ K k, int w, int h)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < w && y < h) k(x, y, (int)blockIdx.z);
}
#endif

#ifndef SS_HOST_EMU
// one CTA of 256 threads per tile; dynamic shared memory sized by the launch
template <class K>
__global__ void __launch_bounds__(256) dm_tile_kernel(const __grid_constant__ K k)
{
    extern __shared__ __align__(16) float dm_tile_smem[];
    k.tile((int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, dm_tile_smem, (int)threadIdx.x, 256);
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
