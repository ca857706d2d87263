// ss_ecc_kernels.cuh -- gradient-based (ECC) estimator of the AlignFrames transform on the GPU.
//
// The reference estimates the frame-to-frame transform from sparse features on the host (SIFT + FLANN +
// RANSAC, reference src/shinestacker/algorithms/align.py:89-148); BASELINE.json's north-star asks for that
// pre-pass as "a CUDA warp-reduction over image gradients" (ECC).  This is that estimator: Evangelidis &
// Psarakis' enhanced-correlation-coefficient iteration (forward additive, the scheme behind
// cv2.findTransformECC) for the 4-parameter similarity that constants.ALIGN_RIGID /
// cv2.estimateAffinePartial2D produce, run coarse to fine on gray pyramids.  Every iteration is ONE pass
// over the reference level: each pixel maps itself into the moving frame, samples value and gradient
// bilinearly, forms its 4 Jacobian entries and contributes to 28 sums (G'G, G'iw, G'ir, G'1, |iw|^2, ir.iw,
// sums and count); warps reduce them with shuffles, CTAs in shared memory, one double atomicAdd per sum and
// CTA.  The 4x4 solve and the ECC bookkeeping run on the host in float64 (ss_ecc_engine.h).
// The reference has no such estimator: parity is "unpinned" and stated through the tolerances of the
// reference's own precision test (tests/test_0031_align_precision.py:62-65) and through agreement with its
// feature pipeline (tests/test_align_ecc_cpu.py, tests/test_gpu_align.py).
#pragma once
#include "ss_dm_kernels.cuh"

#define ECC_NSUMS 28   /* H(10) | G'iw(4) | G'ir(4) | G'1(4) | ww | rw | rr | w | r | n */

// 8-bit gray of a BGR frame as float (utils.img_bw_8bit, reference utils.py:33-42: uint16 frames drop their
// low byte first), cv2's fixed-point BGR2GRAY
template <typename InT>
struct EccGray {
    const void *frame; long long pitch; int h, w; float *out;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const InT *s = (const InT *)((const char *)frame + (long long)y * pitch) + 3 * x;
        const int sh = sizeof(InT) == 2 ? 8 : 0;
        const uint32_t v = (uint32_t)(s[0] >> sh) * 3735u + (uint32_t)(s[1] >> sh) * 19235u +
                           (uint32_t)(s[2] >> sh) * 9798u + 16384u;
        out[(long long)y * w + x] = (float)(v >> 15);
    }
};

// next pyramid level of a gray plane ([1,4,6,4,1]^2 / 256, REFLECT101, as cv2.pyrDown)
struct EccPyrDown {
    const float *in; int h, w; float *out; int h2, w2;
    SS_MDEV void operator()(int x2, int y2, int) const
    {
        int xs[5], ys[5];
        const int x = 2 * x2, y = 2 * y2;
        xs[0] = x; xs[1] = dm_reflect101(x - 1, w); xs[2] = dm_reflect101(x + 1, w);
        xs[3] = dm_reflect101(x - 2, w); xs[4] = dm_reflect101(x + 2, w);
        ys[0] = y; ys[1] = dm_reflect101(y - 1, h); ys[2] = dm_reflect101(y + 1, h);
        ys[3] = dm_reflect101(y - 2, h); ys[4] = dm_reflect101(y + 2, h);
        const float *rows[5];
        for (int i = 0; i < 5; ++i) rows[i] = in + (long long)ys[i] * w;
        out[(long long)y2 * w2 + x2] = dm_pyrdown_px<float>(rows, xs);
    }
};

// central-difference gradient planes of the moving level
struct EccGradient {
    const float *in; int h, w; float *gx, *gy;
    SS_MDEV void operator()(int x, int y, int) const
    {
        const long long o = (long long)y * w + x;
        const int xl = x > 0 ? x - 1 : 0, xr = x + 1 < w ? x + 1 : w - 1;
        const int yu = y > 0 ? y - 1 : 0, yd = y + 1 < h ? y + 1 : h - 1;
        gx[o] = 0.5f * (in[(long long)y * w + xr] - in[(long long)y * w + xl]);
        gy[o] = 0.5f * (in[(long long)yd * w + x] - in[(long long)yu * w + x]);
    }
};

struct EccParams {
    const float *ref, *mov, *gx, *gy; int h, w;
    float a, b, tx, ty;      // reference pixel (x, y) -> moving frame (a x - b y + tx, b x + a y + ty)
    float cx, cy;            // coordinates are taken relative to the level centre (conditioning of G'G)
    float mean_r, mean_w;    // offsets subtracted from the samples before they are summed (less cancellation)
};

SS_DEV float ecc_bilinear(const float *p, int w, int x0, int y0, float fx, float fy)
{
    const float *r0 = p + (long long)y0 * w + x0, *r1 = r0 + w;
    const float top = r0[0] + fx * (r0[1] - r0[0]), bot = r1[0] + fx * (r1[1] - r1[0]);
    return top + fy * (bot - top);
}

// Contribution of one reference pixel to the 28 sums; returns 0 when it maps outside the moving frame.
SS_DEV int ecc_pixel(const EccParams &p, int x, int y, float *s)
{
    const float xc = (float)x - p.cx, yc = (float)y - p.cy;
    const float mx = p.a * xc - p.b * yc + p.tx + p.cx, my = p.b * xc + p.a * yc + p.ty + p.cy;
    if (!(mx >= 0.f && my >= 0.f && mx < (float)(p.w - 1) && my < (float)(p.h - 1))) return 0;
    const int x0 = (int)mx, y0 = (int)my;
    const float fx = mx - (float)x0, fy = my - (float)y0;
    const float iw = ecc_bilinear(p.mov, p.w, x0, y0, fx, fy) - p.mean_w;
    const float gx = ecc_bilinear(p.gx, p.w, x0, y0, fx, fy);
    const float gy = ecc_bilinear(p.gy, p.w, x0, y0, fx, fy);
    const float ir = p.ref[(long long)y * p.w + x] - p.mean_r;
    // d(mx, my)/d(a, b, tx, ty) = [[xc, -yc, 1, 0], [yc, xc, 0, 1]]
    const float g[4] = {gx * xc + gy * yc, gy * xc - gx * yc, gx, gy};
    int k = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) s[k++] = g[i] * g[j];
    for (int i = 0; i < 4; ++i) { s[10 + i] = g[i] * iw; s[14 + i] = g[i] * ir; s[18 + i] = g[i]; }
    s[22] = iw * iw; s[23] = ir * iw; s[24] = ir * ir; s[25] = iw; s[26] = ir; s[27] = 1.f;
    return 1;
}

// ---- the iteration loop of one pyramid level, state kept in device memory ------------------------------
// Everything the host loop of round 1 kept between passes; the float64 algebra on the 28 sums (zero-mean correction,
// 4x4 solve, ECC lambda, step cap, retreat when the correlation falls) runs in ecc_update, on one device thread
// between two grid-wide synchronisations (host build: a plain loop), so a level costs no host round trip at all.
struct EccState {
    double a, b, dx, dy;           // reference -> moving map in absolute coordinates of the current level
    double tx, ty;                 // ... its translation relative to the level centre
    double pa, pb, ptx, pty;       // parameters before the last accepted step
    double step[4];
    double prev_rho, rho, last, overlap, mean_r, mean_w;
    int halvings, done, fail, total, it;
};
enum { ECC_FAIL_NO_OVERLAP = 1, ECC_FAIL_FLAT = 2, ECC_FAIL_SINGULAR = 3 };

// candidate c of the coarse search around (tx, ty): c = (sy + radius) * (2 radius + 1) + (sx + radius)
#ifdef SS_HOST_EMU
inline
#else
__host__ __device__ inline
#endif
EccParams ecc_search_candidate(const EccParams &base, double tx, double ty, int radius, int c)
{
    const int side = 2 * radius + 1;
    EccParams p = base;
    p.tx = (float)(tx + (c % side - radius));
    p.ty = (float)(ty + (c / side - radius));
    return p;
}

struct EccLevelArgs {
    const float *ref, *mov, *gx, *gy;
    int h, w;
    int max_iters;
    int use_state_t;               // 1: tx / ty of the state are valid (set by the coarse search); 0: derive them from dx / dy
    int scale_up;                  // 1: a coarser level than the finest one: dx, dy are doubled for the next level
    double eps;                    // stop when a step moves no pixel by more than this
};

SS_DEV bool ecc_solve4_dev(const double *Hs /*10 upper-triangular*/, const double *b, double *x)
{
    double A[4][5];
    int k = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) { A[i][j] = Hs[k]; A[j][i] = Hs[k]; ++k; }
    for (int i = 0; i < 4; ++i) A[i][4] = b[i];
    for (int c = 0; c < 4; ++c) {            // Gaussian elimination with partial pivoting
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (fabs(A[piv][c]) < 1e-300) return false;
        if (piv != c) for (int j = 0; j < 5; ++j) { const double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        for (int r = c + 1; r < 4; ++r) {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j < 5; ++j) A[r][j] -= f * A[c][j];
        }
    }
    for (int i = 3; i >= 0; --i) {
        double v = A[i][4];
        for (int j = i + 1; j < 4; ++j) v -= A[i][j] * x[j];
        x[i] = v / A[i][i];
    }
    return true;
}

SS_DEV EccParams ecc_params_of(const EccLevelArgs &g, const EccState &st)
{
    EccParams p;
    p.ref = g.ref; p.mov = g.mov; p.gx = g.gx; p.gy = g.gy; p.h = g.h; p.w = g.w;
    p.a = (float)st.a; p.b = (float)st.b; p.tx = (float)st.tx; p.ty = (float)st.ty;
    p.cx = (float)(0.5 * (g.w - 1)); p.cy = (float)(0.5 * (g.h - 1));
    p.mean_r = (float)st.mean_r; p.mean_w = (float)st.mean_w;
    return p;
}

// Start of a level: centre-relative translation, per-level bookkeeping.
SS_DEV void ecc_level_begin(const EccLevelArgs &g, EccState &st)
{
    const double cx = 0.5 * (g.w - 1), cy = 0.5 * (g.h - 1);
    if (!g.use_state_t) {        // mov = A (x - c) + t + c  =>  t = d + A c - c
        st.tx = st.dx + st.a * cx - st.b * cy - cx;
        st.ty = st.dy + st.b * cx + st.a * cy - cy;
    }
    st.prev_rho = -3; st.pa = st.a; st.pb = st.b; st.ptx = st.tx; st.pty = st.ty;
    st.step[0] = st.step[1] = st.step[2] = st.step[3] = 0;
    st.halvings = 0; st.done = 0; st.it = 0; st.mean_r = 0; st.mean_w = 0;
}

// End of a level: back to absolute coordinates (doubled when a finer level follows).
SS_DEV void ecc_level_end(const EccLevelArgs &g, EccState &st)
{
    const double cx = 0.5 * (g.w - 1), cy = 0.5 * (g.h - 1);
    st.dx = st.tx - st.a * cx + st.b * cy + cx;
    st.dy = st.ty - st.b * cx - st.a * cy + cy;
    if (g.scale_up) { st.dx *= 2; st.dy *= 2; }
}

// One iteration's algebra on the sums S of the pass that was just evaluated at the state's parameters.
SS_DEV void ecc_update(const EccLevelArgs &g, EccState &st, const double *S)
{
    st.total++;
    const int it = st.it++;
    const double n = S[27];
    double r_now = -2.0;
    if (n >= 64) {
        const double mw0 = S[25] / n, mr0 = S[26] / n;
        const double sww0 = S[22] - n * mw0 * mw0, srw0 = S[23] - n * mw0 * mr0, srr0 = S[24] - n * mr0 * mr0;
        r_now = sww0 > 0 && srr0 > 0 ? srw0 / sqrt(sww0 * srr0) : -2.0;
    }
    if (it > 0 && r_now < st.prev_rho - 1e-4 && st.halvings < 5) {
        // the last step lowered the correlation (or left the frame): retreat to half of it
        st.halvings++;
        for (int i = 0; i < 4; ++i) st.step[i] *= 0.5;
        st.a = st.pa + st.step[0]; st.b = st.pb + st.step[1]; st.tx = st.ptx + st.step[2]; st.ty = st.pty + st.step[3];
        if (st.it >= g.max_iters) st.done = 1;
        return;
    }
    if (n < 64) { st.fail = ECC_FAIL_NO_OVERLAP; st.done = 1; return; }
    st.halvings = 0;
    st.overlap = n / ((double)g.h * g.w);
    const double mw = S[25] / n, mr = S[26] / n;   // means of the offset values over the overlap
    double Gw[4], Gr[4];
    for (int i = 0; i < 4; ++i) { Gw[i] = S[10 + i] - mw * S[18 + i]; Gr[i] = S[14 + i] - mr * S[18 + i]; }
    const double sww = S[22] - n * mw * mw, srw = S[23] - n * mw * mr, srr = S[24] - n * mr * mr;
    if (sww <= 0 || srr <= 0) { st.fail = ECC_FAIL_FLAT; st.done = 1; return; }
    st.rho = srw / sqrt(sww * srr);
    st.prev_rho = st.rho; st.pa = st.a; st.pb = st.b; st.ptx = st.tx; st.pty = st.ty;
    double HiGw[4];
    if (!ecc_solve4_dev(S, Gw, HiGw)) { st.fail = ECC_FAIL_SINGULAR; st.done = 1; return; }
    double num = sww, den = srw;
    for (int i = 0; i < 4; ++i) { num -= Gw[i] * HiGw[i]; den -= Gr[i] * HiGw[i]; }
    const double lambda = den > 0 ? num / den : (srw > 0 ? sww / srw : 1.0);
    double rhs[4], dp[4];
    for (int i = 0; i < 4; ++i) rhs[i] = lambda * Gr[i] - Gw[i];
    if (!ecc_solve4_dev(S, rhs, dp)) { st.fail = ECC_FAIL_SINGULAR; st.done = 1; return; }
    const double radius = 0.5 * sqrt((double)g.w * g.w + (double)g.h * g.h);
    st.last = sqrt(dp[2] * dp[2] + dp[3] * dp[3]) + radius * sqrt(dp[0] * dp[0] + dp[1] * dp[1]);
    const double cap = st.last > 2.0 ? 2.0 / st.last : 1.0;   // no pixel moves by more than ~2 px per step
    for (int i = 0; i < 4; ++i) st.step[i] = dp[i] * cap;
    st.a += st.step[0]; st.b += st.step[1]; st.tx += st.step[2]; st.ty += st.step[3];
    // the offsets only condition the float sums; keep them fixed after the first pass of a level
    if (it == 0) { st.mean_r += mr; st.mean_w += mw; }
    if (st.last < g.eps || st.it >= g.max_iters) st.done = 1;
}

#ifndef SS_HOST_EMU
// All iterations of one level in ONE cooperative launch: every pass is a grid-stride accumulation into one of two sum
// buffers, a grid-wide synchronisation, the update on one thread, a second synchronisation.
__global__ void __launch_bounds__(256) ecc_level_kernel(const __grid_constant__ EccLevelArgs g, EccState *st, double *sums,
                                                        unsigned int *bar)
{
    __shared__ float part[8][ECC_NSUMS];
    __shared__ int s_done;
    const unsigned int nblk = gridDim.x;
    unsigned int epoch = 0;
    // grid-wide barrier on a monotonically increasing counter (all CTAs of the launch are resident: cooperative launch)
    auto grid_sync = [&]() {
        __syncthreads();
        if (threadIdx.x == 0) {
            ++epoch;
            __threadfence();
            atomicAdd(bar, 1u);
            while (*(volatile unsigned int *)bar < epoch * nblk) { }
            __threadfence();
        }
        __syncthreads();
    };
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (!st->fail) ecc_level_begin(g, *st); else st->done = 1;
        for (int i = 0; i < 2 * ECC_NSUMS; ++i) sums[i] = 0.0;
    }
    grid_sync();
    const long long n = (long long)g.h * g.w;
    for (int it = 0; it < g.max_iters + 8; ++it) {
        if (threadIdx.x == 0) s_done = ((volatile EccState *)st)->done;
        __syncthreads();
        if (s_done) break;
        EccParams p;      // the state was written by another SM: read it around the L1 cache
        {
            const volatile EccState *vs = st;
            p.ref = g.ref; p.mov = g.mov; p.gx = g.gx; p.gy = g.gy; p.h = g.h; p.w = g.w;
            p.a = (float)vs->a; p.b = (float)vs->b; p.tx = (float)vs->tx; p.ty = (float)vs->ty;
            p.cx = (float)(0.5 * (g.w - 1)); p.cy = (float)(0.5 * (g.h - 1));
            p.mean_r = (float)vs->mean_r; p.mean_w = (float)vs->mean_w;
        }
        double *out = sums + (it & 1) * ECC_NSUMS;
        float acc[ECC_NSUMS];
#pragma unroll
        for (int i = 0; i < ECC_NSUMS; ++i) acc[i] = 0.f;
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
            float s[ECC_NSUMS];
            if (ecc_pixel(p, (int)(idx % p.w), (int)(idx / p.w), s)) {
#pragma unroll
                for (int i = 0; i < ECC_NSUMS; ++i) acc[i] += s[i];
            }
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < ECC_NSUMS; ++i) {
            float v = acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) part[warp][i] = v;
        }
        __syncthreads();
        if (threadIdx.x < ECC_NSUMS) {
            double v = 0.0;
            for (int wi = 0; wi < 8; ++wi) v += (double)part[wi][threadIdx.x];
            atomicAdd(out + threadIdx.x, v);
        }
        grid_sync();
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            double S[ECC_NSUMS];
            for (int i = 0; i < ECC_NSUMS; ++i) S[i] = ((const volatile double *)out)[i];
            double *other = sums + ((it + 1) & 1) * ECC_NSUMS;
            for (int i = 0; i < ECC_NSUMS; ++i) other[i] = 0.0;
            ecc_update(g, *st, S);
        }
        grid_sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && !st->fail) ecc_level_end(g, *st);
    // (the barrier counter is zeroed by the host before every launch)
}

// the state as a kernel argument: no host-to-device copy (which would queue behind the frame uploads of the copy engine)
__global__ void ecc_set_state_kernel(const __grid_constant__ EccState st, EccState *dst) { *dst = st; }

// One ECC iteration over a level: grid-stride over the pixels, per-thread float partial sums, shuffle
// reduction per warp, shared-memory reduction per CTA, one double atomicAdd per sum and CTA.
// blockIdx.y selects the candidate of the coarse search (one launch for all): whole-pixel shifts (sx, sy) in
// [-radius, radius]^2 around (tx, ty), sy outermost -- ecc_search_candidate().
__global__ void __launch_bounds__(256) ecc_accumulate_kernel(const __grid_constant__ EccParams base, double tx, double ty,
                                                             int radius, double *all_sums)
{
    const EccParams p = ecc_search_candidate(base, tx, ty, radius, (int)blockIdx.y);
    double *sums = all_sums + (size_t)blockIdx.y * ECC_NSUMS;
    float acc[ECC_NSUMS];
#pragma unroll
    for (int i = 0; i < ECC_NSUMS; ++i) acc[i] = 0.f;
    const long long n = (long long)p.h * p.w;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        float s[ECC_NSUMS];
        if (ecc_pixel(p, (int)(idx % p.w), (int)(idx / p.w), s)) {
#pragma unroll
            for (int i = 0; i < ECC_NSUMS; ++i) acc[i] += s[i];
        }
    }
    __shared__ float part[8][ECC_NSUMS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < ECC_NSUMS; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) part[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < ECC_NSUMS) {
        double v = 0.0;
        for (int wi = 0; wi < 8; ++wi) v += (double)part[wi][threadIdx.x];
        atomicAdd(sums + threadIdx.x, v);
    }
}
#endif
