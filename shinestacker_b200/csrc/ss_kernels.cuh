// ss_kernels.cuh -- device code of the B200 focus-stacking core (sm_100a).
//
// One body, two builds: nvcc compiles it as CUDA kernels (the product); tests/emu/ compiles the very
// same tile code with g++ (-DSS_HOST_EMU) so the index arithmetic (halos, REFLECT101 fix-ups, tail
// columns) can be checked on a CPU-only box.  The emulation is test infrastructure; the product never
// uses it.
//
// Arithmetic contract (what makes the result bit-identical to the reference CPU path):
//   * every 5x5 filter output is ONE sequential chain over the 25 taps in row-major order starting
//     from +0: fused multiply-add for output columns x < w - w%8, separate multiply and add for the
//     last w%8 columns (w = width of the array the reference hands to cv2.filter2D);
//   * expand skips the taps that fall on stuffed zeros (exact), keeps row-major order, then *4;
//   * gray = fma(R,.299f, fma(B,.114f, G*.587f));  energy = filter(gray*gray);
//   * arg-max over frames = strict '>' in frame order; winner coefficients are stored as 0 + lap.
// Build with -fmad=false: every FMA below is explicit and none may be created by the compiler.
#pragma once
#include <stdint.h>

#ifdef SS_HOST_EMU
#include <math.h>
#define SS_DEV static inline
#define SS_THREADS(tid) for (int tid = 0; tid < SS_NT; ++tid)
#define SS_SYNC()
#define SS_PT_DECL(type, name, n) type name[SS_NT][n]
#define SS_PT(name) name[tid]
SS_DEV float ss_fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
SS_DEV float ss_mul(float a, float b) { return a * b; }
SS_DEV float ss_add(float a, float b) { return a + b; }
SS_DEV float ss_sub(float a, float b) { return a - b; }
#else
#define SS_DEV __device__ __forceinline__
#define SS_THREADS(tid) for (int tid = threadIdx.x, ss_once_ = 1; ss_once_; ss_once_ = 0)
#define SS_SYNC() __syncthreads()
#define SS_PT_DECL(type, name, n) type name[n]
#define SS_PT(name) name
SS_DEV float ss_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SS_DEV float ss_mul(float a, float b) { return __fmul_rn(a, b); }
SS_DEV float ss_add(float a, float b) { return __fadd_rn(a, b); }
SS_DEV float ss_sub(float a, float b) { return __fsub_rn(a, b); }
#endif

// ---- tile geometry of the fused level kernel --------------------------------------------------------
#define SS_NT 256            // threads per CTA
#define SS_TX 64             // output tile width  (pixels of level l)
#define SS_TY 32             // output tile height
#define SS_P0W (SS_TX + 12)  // g_l patch: halo 6  (energy 2 + expand 2 (in g_{l+1} units 1) + reduce 2)
#define SS_P0H (SS_TY + 12)
#define SS_P0S 77            // row stride of the g_l patch in floats (odd: fewer bank conflicts)
#define SS_P1W (SS_TX / 2 + 4)  // g_{l+1} patch
#define SS_P1H (SS_TY / 2 + 4)
#define SS_P1S 37
#define SS_SQW (SS_TX + 4)   // gray^2 patch: halo 2
#define SS_SQH (SS_TY + 4)
#define SS_SQS 69
#define SS_LPS 65            // winner-candidate Laplacian tile row stride
#define SS_SMEM_FLOATS (3 * SS_P0H * SS_P0S + 3 * SS_P1H * SS_P1S + SS_SQH * SS_SQS + 3 * SS_TY * SS_LPS)

struct SsLevelParams {
    // level l geometry
    int h, w;            // size of g_l
    int h2, w2;          // size of g_{l+1} = ceil(h/2), ceil(w/2)
    // input g_l: level 0 -> interleaved BGR samples (uint8/uint16), else planar float32
    const void *in;      // first frame of the batch
    long long in_frame_stride;   // bytes between frames
    long long in_plane_stride;   // planar only: bytes between channel planes
    int in_pitch;                // bytes per row
    // output g_{l+1}: planar float32, one slot per frame of the batch
    float *gnext;
    long long gnext_frame_stride;  // floats
    long long gnext_plane_stride;  // floats
    int gnext_pitch;               // floats per row
    // running state of level l (planar, tightly packed h x w)
    float *best_e;
    int32_t *best_idx;
    float *fused;        // 3 planes, plane stride h*w
    int n_frames;        // frames in this batch
    int frame_index0;    // global index of the first frame of the batch
    int init;            // 1: the batch starts the stack (no prior state)
    float k25[25];
};

SS_DEV int ss_reflect_clamp(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    if (i < 0) i = 0;
    if (i >= n) i = n - 1;
    return i;
}

template <bool TAIL>
SS_DEV float ss_mac(float v, float k, float acc)
{
    return TAIL ? ss_add(acc, ss_mul(v, k)) : ss_fma(v, k, acc);
}

template <typename InT>
SS_DEV void ss_load_px(const SsLevelParams &p, const char *frame, int y, int x, float &b, float &g,
                       float &r)
{
    if constexpr (sizeof(InT) == 4) {  // planar float32 (levels >= 1)
        const char *s = frame + (long long)y * p.in_pitch + 4ll * x;
        b = *(const float *)s;
        g = *(const float *)(s + p.in_plane_stride);
        r = *(const float *)(s + 2 * p.in_plane_stride);
    } else {                           // interleaved BGR uint8 / uint16 (level 0)
        const InT *s = (const InT *)(frame + (long long)y * p.in_pitch) + 3 * x;
        b = (float)s[0]; g = (float)s[1]; r = (float)s[2];
    }
}

// 25-tap chain on a patch with row stride S, taps at rows r0..r0+4, columns c0..c0+4.
template <bool TAIL, int S>
SS_DEV float ss_conv25(const float *patch, int r0, int c0, const float *k25)
{
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx)
            acc = ss_mac<TAIL>(patch[(r0 + dy) * S + c0 + dx], k25[dy * 5 + dx], acc);
    return acc;
}

// The fused level step for one tile: for every frame of the batch
//   S0  g_l patch (halo 6, REFLECT101 at the image border)              -> smem P0
//   S1  g_{l+1} = reduce(g_l) on the patch (reference pyramid.py:27-32) -> smem P1 + global (interior)
//   S2  lap = g_l - 4*expand(g_{l+1}) (pyramid.py:34-46,133-138), gray^2 (pyramid.py:49-50) -> smem SQ, LAP
//   S3  energy = filter(gray^2); running strict-'>' arg-max (pyramid.py:50-54) in registers
// and finally the running state goes back to global memory.
template <typename InT>
#ifndef SS_HOST_EMU
__device__ __forceinline__
#else
static inline
#endif
void ss_level_tile(const SsLevelParams &p, int tile_x, int tile_y, float *smem)
{
    float *P0 = smem;
    float *P1 = P0 + 3 * SS_P0H * SS_P0S;
    float *SQ = P1 + 3 * SS_P1H * SS_P1S;
    float *LAP = SQ + SS_SQH * SS_SQS;
    const int x0 = tile_x * SS_TX, y0 = tile_y * SS_TY;
    const int w_vec = p.w - (p.w % 8);              // FMA columns of the reduce / energy filters
    const int wz = 2 * p.w2, wz_vec = wz - (wz % 8);  // ... of the expand filter (width 2*w2)
    const int j0 = x0 / 2 - 2, i0 = y0 / 2 - 2;       // g_{l+1} coordinates of P1[0][0]
    // does any patch position of this tile fall outside the images?
    const bool border = (x0 < 6) || (y0 < 6) || (x0 + SS_TX + 6 > p.w) || (y0 + SS_TY + 6 > p.h);

    SS_PT_DECL(float, be, 8);
    SS_PT_DECL(int, bi, 8);
    SS_PT_DECL(float, bl, 24);

    // thread -> owned 4x2 pixel block
    if (!p.init) {
        SS_THREADS(tid)
        {
            const int tx = tid & 15, ty = tid >> 4;
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 4; ++c) {
                    const int x = x0 + 4 * tx + c, y = y0 + 2 * ty + r, q = r * 4 + c;
                    if (x < p.w && y < p.h) {
                        const long long o = (long long)y * p.w + x;
                        SS_PT(be)[q] = p.best_e[o];
                        SS_PT(bi)[q] = p.best_idx[o];
                        for (int ch = 0; ch < 3; ++ch)
                            SS_PT(bl)[ch * 8 + q] = p.fused[(long long)ch * p.h * p.w + o];
                    }
                }
        }
    }

    for (int f = 0; f < p.n_frames; ++f) {
        const char *frame = (const char *)p.in + (long long)f * p.in_frame_stride;
        // ---- S0: g_l patch -------------------------------------------------------------------------
        SS_THREADS(tid)
        {
            for (int i = tid; i < SS_P0H * SS_P0W; i += SS_NT) {
                const int r = i / SS_P0W, c = i - r * SS_P0W;
                const int y = ss_reflect_clamp(y0 - 6 + r, p.h), x = ss_reflect_clamp(x0 - 6 + c, p.w);
                float b, g, rr;
                ss_load_px<InT>(p, frame, y, x, b, g, rr);
                P0[(0 * SS_P0H + r) * SS_P0S + c] = b;
                P0[(1 * SS_P0H + r) * SS_P0S + c] = g;
                P0[(2 * SS_P0H + r) * SS_P0S + c] = rr;
            }
        }
        SS_SYNC();
        // ---- S1: reduce -> P1 (in-image positions), interior -> global -----------------------------
        SS_THREADS(tid)
        {
            for (int i = tid; i < 3 * SS_P1H * SS_P1W; i += SS_NT) {
                const int ch = i / (SS_P1H * SS_P1W), rem = i - ch * (SS_P1H * SS_P1W);
                const int ii = rem / SS_P1W, jj = rem - ii * SS_P1W;
                const int I = i0 + ii, J = j0 + jj;
                if (I < 0 || I >= p.h2 || J < 0 || J >= p.w2) continue;
                const float *patch = P0 + ch * SS_P0H * SS_P0S;
                float v;
                if (2 * J >= w_vec) v = ss_conv25<true, SS_P0S>(patch, 2 * ii, 2 * jj, p.k25);
                else v = ss_conv25<false, SS_P0S>(patch, 2 * ii, 2 * jj, p.k25);
                P1[(ch * SS_P1H + ii) * SS_P1S + jj] = v;
                if (ii >= 2 && ii < SS_P1H - 2 && jj >= 2 && jj < SS_P1W - 2)
                    p.gnext[f * p.gnext_frame_stride + ch * p.gnext_plane_stride +
                            (long long)I * p.gnext_pitch + J] = v;
            }
        }
        SS_SYNC();
        if (border) {
            // out-of-image g_{l+1} positions as the expand filter sees them: the stuffed image has size
            // 2*h2 x 2*w2 and REFLECT101 on it maps even index -2 -> 2 (J=-1 -> 1) and 2*w2 -> 2*w2-2
            // (J=w2 -> w2-1).
            SS_THREADS(tid)
            {
                for (int i = tid; i < 3 * SS_P1H * SS_P1W; i += SS_NT) {
                    const int ch = i / (SS_P1H * SS_P1W), rem = i - ch * (SS_P1H * SS_P1W);
                    const int ii = rem / SS_P1W, jj = rem - ii * SS_P1W;
                    const int I = i0 + ii, J = j0 + jj;
                    if (I >= 0 && I < p.h2 && J >= 0 && J < p.w2) continue;
                    const int Is = I < 0 ? -I : (I >= p.h2 ? 2 * p.h2 - 1 - I : I);
                    const int Js = J < 0 ? -J : (J >= p.w2 ? 2 * p.w2 - 1 - J : J);
                    if (Is < 0 || Is >= p.h2 || Js < 0 || Js >= p.w2) continue;  // only junk needs it
                    const int is = Is - i0, js = Js - j0;
                    if (is < 0 || is >= SS_P1H || js < 0 || js >= SS_P1W) continue;
                    P1[(ch * SS_P1H + ii) * SS_P1S + jj] = P1[(ch * SS_P1H + is) * SS_P1S + js];
                }
            }
            SS_SYNC();
        }
        // ---- S2: Laplacian, gray, square -----------------------------------------------------------
        SS_THREADS(tid)
        {
            for (int i = tid; i < SS_SQH * SS_SQW; i += SS_NT) {
                const int qq = i / SS_SQW, pp = i - qq * SS_SQW;
                const int y = y0 - 2 + qq, x = x0 - 2 + pp;
                if (y < 0 || y >= p.h || x < 0 || x >= p.w) continue;
                const bool tail = x >= wz_vec;
                float lap[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float *g1 = P1 + ch * SS_P1H * SS_P1S;
                    float acc = 0.f;
#pragma unroll
                    for (int dy = -2; dy <= 2; ++dy) {
                        const int zy = y + dy;
                        if (zy & 1) continue;
                        const int ii = (zy >> 1) - i0;
#pragma unroll
                        for (int dx = -2; dx <= 2; ++dx) {
                            const int zx = x + dx;
                            if (zx & 1) continue;
                            const int jj = (zx >> 1) - j0;
                            const float v = g1[ii * SS_P1S + jj], k = p.k25[(dy + 2) * 5 + dx + 2];
                            acc = tail ? ss_mac<true>(v, k, acc) : ss_mac<false>(v, k, acc);
                        }
                    }
                    const float e = ss_mul(4.f, acc);
                    lap[ch] = ss_sub(P0[(ch * SS_P0H + qq + 4) * SS_P0S + pp + 4], e);
                }
                const float gray = ss_fma(lap[2], 0.299f, ss_fma(lap[0], 0.114f, ss_mul(lap[1], 0.587f)));
                SQ[qq * SS_SQS + pp] = ss_mul(gray, gray);
                if (qq >= 2 && qq < SS_SQH - 2 && pp >= 2 && pp < SS_SQW - 2) {
                    const int o = (qq - 2) * SS_LPS + pp - 2;
                    LAP[o] = lap[0];
                    LAP[SS_TY * SS_LPS + o] = lap[1];
                    LAP[2 * SS_TY * SS_LPS + o] = lap[2];
                }
            }
        }
        SS_SYNC();
        if (border) {
            // out-of-image gray^2 positions: REFLECT101 of the energy filter at the size of g_l
            SS_THREADS(tid)
            {
                for (int i = tid; i < SS_SQH * SS_SQW; i += SS_NT) {
                    const int qq = i / SS_SQW, pp = i - qq * SS_SQW;
                    const int y = y0 - 2 + qq, x = x0 - 2 + pp;
                    if (y >= 0 && y < p.h && x >= 0 && x < p.w) continue;
                    const int ys = y < 0 ? -y : (y >= p.h ? 2 * p.h - 2 - y : y);
                    const int xs = x < 0 ? -x : (x >= p.w ? 2 * p.w - 2 - x : x);
                    if (ys < 0 || ys >= p.h || xs < 0 || xs >= p.w) continue;
                    const int qs = ys - (y0 - 2), ps = xs - (x0 - 2);
                    if (qs < 0 || qs >= SS_SQH || ps < 0 || ps >= SS_SQW) continue;
                    SQ[qq * SS_SQS + pp] = SQ[qs * SS_SQS + ps];
                }
            }
            SS_SYNC();
        }
        // ---- S3: energy + running arg-max ----------------------------------------------------------
        SS_THREADS(tid)
        {
            const int tx = tid & 15, ty = tid >> 4;
            const int gidx = p.frame_index0 + f;
            const bool take_all = p.init && f == 0;
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int xx = 4 * tx + c, yy = 2 * ty + r, q = r * 4 + c;
                    const int x = x0 + xx, y = y0 + yy;
                    if (x >= p.w || y >= p.h) continue;
                    float e;
                    if (x >= w_vec) e = ss_conv25<true, SS_SQS>(SQ, yy, xx, p.k25);
                    else e = ss_conv25<false, SS_SQS>(SQ, yy, xx, p.k25);
                    if (take_all || e > SS_PT(be)[q]) {
                        SS_PT(be)[q] = e;
                        SS_PT(bi)[q] = gidx;
                        const int o = yy * SS_LPS + xx;
                        SS_PT(bl)[q] = ss_add(0.f, LAP[o]);
                        SS_PT(bl)[8 + q] = ss_add(0.f, LAP[SS_TY * SS_LPS + o]);
                        SS_PT(bl)[16 + q] = ss_add(0.f, LAP[2 * SS_TY * SS_LPS + o]);
                    }
                }
        }
        // no barrier needed here: the next frame's S0/S1 only overwrite P0/P1 (last read in S2, two
        // barriers ago); SQ/LAP are rewritten in S2, two barriers after this S3.
    }

    SS_THREADS(tid)
    {
        const int tx = tid & 15, ty = tid >> 4;
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 4; ++c) {
                const int x = x0 + 4 * tx + c, y = y0 + 2 * ty + r, q = r * 4 + c;
                if (x < p.w && y < p.h) {
                    const long long o = (long long)y * p.w + x;
                    p.best_e[o] = SS_PT(be)[q];
                    p.best_idx[o] = SS_PT(bi)[q];
                    for (int ch = 0; ch < 3; ++ch)
                        p.fused[(long long)ch * p.h * p.w + o] = SS_PT(bl)[ch * 8 + q];
                }
            }
    }
}
