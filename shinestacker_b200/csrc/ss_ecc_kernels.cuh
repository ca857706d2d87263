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

#ifndef SS_HOST_EMU
// One ECC iteration over a level: grid-stride over the pixels, per-thread float partial sums, shuffle
// reduction per warp, shared-memory reduction per CTA, one double atomicAdd per sum and CTA.
__global__ void __launch_bounds__(256) ecc_accumulate_kernel(const __grid_constant__ EccParams p, double *sums)
{
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
