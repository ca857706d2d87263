// ss_al_kernels.cuh -- device code of the warp stage of AlignFrames (sm_100a): what the reference's
// align_images does once the transform is known (reference src/shinestacker/algorithms/align.py:216-251):
// cv2.warpAffine / cv2.warpPerspective (INTER_LINEAR; constant or replicated border), the validity mask of
// BORDER_REPLICATE_BLUR, cv2.GaussianBlur(21x21) and the paste of the blurred pixels where the mask is 0.
// Pixel functors like ss_dm_kernels.cuh: nvcc instantiates dm_px_kernel<Functor>, tests/emu/ runs the same
// functors in loops under g++.
//
// All of it is integer / fixed-point work and bit-exact against cv2 4.13 (oracle/align_oracle.py):
//   * source coordinates in 1/32 pixel: affine from the host-built tables rint(M*x*1024) (+16, >> 5) --
//     float64 only on the host --, perspective rint(32 * X/W) in float64 per pixel;
//   * uint8: (sum_i w_i p_i * 32 + 2^14) >> 15 with w = (32-fx)(32-fy), ...; uint16: float32
//     v0*w0 + v1*w1 + v2*w2 + v3*w3 (unfused, left to right), round-half-even;
//   * mask != 0 <=> the taps inside the image carry >= 512/1024 of the weight;
//   * blur: (sum_ij k_i k_j p_ij + half) >> shift with cv2's integer kernel (8 / 16 fractional bits),
//     evaluated only where the mask is 0 (thin border strips), REFLECT101.
// uint16 + homography follows cv2 4.13's float kernels instead (not bit-pinned, see the oracle's header).
#pragma once
#include "ss_dm_kernels.cuh"

#define AL_MAX_BLUR_TAPS 64

#ifdef SS_HOST_EMU
#include <math.h>
SS_DEV long long al_rint_ll(double v) { return (long long)nearbyint(v); }
SS_DEV float al_rintf(float v) { return nearbyintf(v); }
#else
SS_DEV long long al_rint_ll(double v) { return __double2ll_rn(v); }
SS_DEV float al_rintf(float v) { return rintf(v); }
#endif

// list of the pixels the blur has to fill (mask == 0): appended by AlWarp, consumed one warp per pixel
SS_DEV void al_append(int *counter, int *list, int idx)
{
#ifdef SS_HOST_EMU
    list[(*counter)++] = idx;
#else
    list[atomicAdd(counter, 1)] = idx;
#endif
}
SS_DEV unsigned long long al_lane_sum(unsigned long long v)
{
#ifndef SS_HOST_EMU
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
}

template <typename T>
struct AlWarp {
    const T *src; long long spitch; int h, w;
    int persp;                                   // 0: affine tables, 1: perspective (inverted matrix in m)
    const int *adelta, *bdelta, *x0, *y0;        // affine: rint(M0*x*1024), rint(M3*x*1024), row offsets + 16
    double m[9];
    int border_const; int bval[3];               // BORDER_CONSTANT and its value (already saturated to T)
    int float_persp;                             // uint16 + homography: cv2's float path
    T *dst; long long dpitch;
    T *dst2;                                     // second copy of the result (nullable): the buffer the blur fill patches
    int *fill_count, *fill_list;                 // nullable: pixels whose mask (warp of ones, constant 0 border) is 0

    SS_MDEV const T *row(int y) const { return (const T *)((const char *)src + (long long)y * spitch); }
    SS_MDEV void operator()(int x, int y, int) const
    {
        long long X, Y;
        float ffx = 0.f, ffy = 0.f;
        long long isx = 0, isy = 0;             // float path: integer source position (floor of the exact quotient)
        if (!persp) {
            X = ((long long)x0[y] + adelta[x]) >> 5;
            Y = ((long long)y0[y] + bdelta[x]) >> 5;
        } else {
            const double dx = (double)x, dy = (double)y;
            const double X0 = m[0] * dx + (m[1] * dy + m[2]);
            const double Y0 = m[3] * dx + (m[4] * dy + m[5]);
            double W = m[6] * dx + (m[7] * dy + m[8]);
            if (float_persp) {
                // cv2 4.13, uint16 + warpPerspective: quotient and its fraction in float64, the fraction then as float32
                double qx = X0 / W, qy = Y0 / W;
                qx = qx > -1e9 ? (qx < 1e9 ? qx : 1e9) : -1e9;       // also catches NaN (W = 0)
                qy = qy > -1e9 ? (qy < 1e9 ? qy : 1e9) : -1e9;
                const double flx = floor(qx), fly = floor(qy);
                ffx = (float)(qx - flx); ffy = (float)(qy - fly);
                isx = (long long)flx; isy = (long long)fly;
            }
            W = W != 0.0 ? 32.0 / W : 0.0;      // the 1/32-pixel coordinates: every other type, and the mask of this one
            double fX = X0 * W, fY = Y0 * W;
            fX = fX < -2147483648.0 ? -2147483648.0 : (fX > 2147483647.0 ? 2147483647.0 : fX);
            fY = fY < -2147483648.0 ? -2147483648.0 : (fY > 2147483647.0 ? 2147483647.0 : fY);
            X = al_rint_ll(fX); Y = al_rint_ll(fY);
        }
        if (!float_persp) {
            // interior fast path: all four taps inside the image (the overwhelming majority of the pixels)
            const long long sxi = X >> 5, syi = Y >> 5;
            if (sxi >= 0 && sxi < w - 1 && syi >= 0 && syi < h - 1) {
                const int fx = (int)(X & 31), fy = (int)(Y & 31);
                const int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
                const T *p0 = row((int)syi) + 3 * (int)sxi;
                const T *p1 = (const T *)((const char *)p0 + spitch);
                T *o = (T *)((char *)dst + (long long)y * dpitch) + 3 * x;
                T res[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int v00 = p0[c], v01 = p0[3 + c], v10 = p1[c], v11 = p1[3 + c];
                    if (sizeof(T) == 1) {
                        res[c] = (T)(((v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11) * 32 + 16384) >> 15);
                    } else {
                        float t = ss_mul((float)v00, ss_mul((float)w00, 0.0009765625f));
                        t = ss_add(t, ss_mul((float)v01, ss_mul((float)w01, 0.0009765625f)));
                        t = ss_add(t, ss_mul((float)v10, ss_mul((float)w10, 0.0009765625f)));
                        t = ss_add(t, ss_mul((float)v11, ss_mul((float)w11, 0.0009765625f)));
                        t = al_rintf(t);
                        res[c] = (T)(int)(t > 65535.f ? 65535.f : t);
                    }
                }
                o[0] = res[0]; o[1] = res[1]; o[2] = res[2];
                if (dst2) {
                    T *o2 = (T *)((char *)dst2 + (long long)y * dpitch) + 3 * x;
                    o2[0] = res[0]; o2[1] = res[1]; o2[2] = res[2];
                }
                return;
            }
        }
        long long sxl = X >> 5, syl = Y >> 5;
        sxl = sxl < -32768 ? -32768 : (sxl > 32767 ? 32767 : sxl);
        syl = syl < -32768 ? -32768 : (syl > 32767 ? 32767 : syl);
        const int sx = (int)sxl, sy = (int)syl, fx = (int)(X & 31), fy = (int)(Y & 31);
        const int wt[4] = {(32 - fx) * (32 - fy), fx * (32 - fy), (32 - fx) * fy, fx * fy};
        const int in_x0 = sx >= 0 && sx < w, in_x1 = sx + 1 >= 0 && sx + 1 < w;
        const int in_y0 = sy >= 0 && sy < h, in_y1 = sy + 1 >= 0 && sy + 1 < h;
        const int inside[4] = {in_x0 && in_y0, in_x1 && in_y0, in_x0 && in_y1, in_x1 && in_y1};
        const int cx0 = sx < 0 ? 0 : (sx > w - 1 ? w - 1 : sx), cx1 = sx + 1 < 0 ? 0 : (sx + 1 > w - 1 ? w - 1 : sx + 1);
        const int cy0 = sy < 0 ? 0 : (sy > h - 1 ? h - 1 : sy), cy1 = sy + 1 < 0 ? 0 : (sy + 1 > h - 1 ? h - 1 : sy + 1);
        T *o = (T *)((char *)dst + (long long)y * dpitch) + 3 * x;
        if (fill_list) {
            int s = 0;
            for (int i = 0; i < 4; ++i) s += inside[i] ? wt[i] : 0;
            if (s < 512) al_append(fill_count, fill_list, y * w + x);
        }
        if (float_persp) {
            // vertical interpolation first, each step one fma, round half to even: bit-identical to cv2 4.13 wherever all
            // four taps lie in the image.  With a tap outside, cv2 writes the border VALUE (BORDER_CONSTANT) or nothing at
            // all (BORDER_REPLICATE: the destination keeps what it held); here REPLICATE replicates the edge pixel.
            const long long lx = isx < -32768 ? -32768 : (isx > 32767 ? 32767 : isx), ly = isy < -32768 ? -32768 : (isy > 32767 ? 32767 : isy);
            const int tx = (int)lx, ty = (int)ly;
            const bool all_in = tx >= 0 && tx + 1 < w && ty >= 0 && ty + 1 < h;
            const int ax0 = tx < 0 ? 0 : (tx > w - 1 ? w - 1 : tx), ax1 = tx + 1 < 0 ? 0 : (tx + 1 > w - 1 ? w - 1 : tx + 1);
            const int ay0 = ty < 0 ? 0 : (ty > h - 1 ? h - 1 : ty), ay1 = ty + 1 < 0 ? 0 : (ty + 1 > h - 1 ? h - 1 : ty + 1);
            const T *q0 = row(ay0), *q1 = row(ay1);
            for (int c = 0; c < 3; ++c) {
                if (border_const && !all_in) { o[c] = (T)bval[c]; continue; }
                const float v00 = (float)q0[3 * ax0 + c], v01 = (float)q0[3 * ax1 + c];
                const float v10 = (float)q1[3 * ax0 + c], v11 = (float)q1[3 * ax1 + c];
                const float l = ss_fma(ffy, ss_add(v10, -v00), v00), r = ss_fma(ffy, ss_add(v11, -v01), v01);
                float t = al_rintf(ss_fma(ffx, ss_add(r, -l), l));
                t = t < 0.f ? 0.f : (t > 65535.f ? 65535.f : t);
                o[c] = (T)(int)t;
            }
            if (dst2) {
                T *o2 = (T *)((char *)dst2 + (long long)y * dpitch) + 3 * x;
                o2[0] = o[0]; o2[1] = o[1]; o2[2] = o[2];
            }
            return;
        }
        const T *r0 = row(cy0), *r1 = row(cy1);
        for (int c = 0; c < 3; ++c) {
            int v[4] = {(int)r0[3 * cx0 + c], (int)r0[3 * cx1 + c], (int)r1[3 * cx0 + c], (int)r1[3 * cx1 + c]};
            if (border_const)
                for (int i = 0; i < 4; ++i) if (!inside[i]) v[i] = bval[c];
            if (sizeof(T) == 1) {
                const int t = ((v[0] * wt[0] + v[1] * wt[1] + v[2] * wt[2] + v[3] * wt[3]) * 32 + 16384) >> 15;
                o[c] = (T)(t > 255 ? 255 : t);
            } else {
                float t = ss_mul((float)v[0], ss_mul((float)wt[0], 0.0009765625f));
                t = ss_add(t, ss_mul((float)v[1], ss_mul((float)wt[1], 0.0009765625f)));
                t = ss_add(t, ss_mul((float)v[2], ss_mul((float)wt[2], 0.0009765625f)));
                t = ss_add(t, ss_mul((float)v[3], ss_mul((float)wt[3], 0.0009765625f)));
                t = al_rintf(t);
                o[c] = (T)(int)(t > 65535.f ? 65535.f : t);
            }
        }
        if (dst2) {
            T *o2 = (T *)((char *)dst2 + (long long)y * dpitch) + 3 * x;
            o2[0] = o[0]; o2[1] = o[1]; o2[2] = o[2];
        }
    }
};

// cv2.warpAffine's fixed-point coordinate tables (AB_BITS = 10) for the inverted 2x3 matrix M, built on the device
// (a host-built table would need a host-to-device copy that queues behind the frame uploads): one call = entry i of
// the column tables (i < w) and of the row tables (i < h).  Float64 products and sums unfused, round half to even:
// the values of cv2's saturate_cast<int>(M * x * 1024).
struct AlAffineTables {
    double M[6]; int w, h; int *ad, *bd, *x0, *y0;
    SS_MDEV void operator()(int i, int, int) const
    {
        const double v = (double)i;
        if (i < w) {
            ad[i] = (int)al_rint_ll(M[0] * v * 1024);
            bd[i] = (int)al_rint_ll(M[3] * v * 1024);
        }
        if (i < h) {
            x0[i] = (int)al_rint_ll((M[1] * v + M[2]) * 1024) + 16;
            y0[i] = (int)al_rint_ll((M[4] * v + M[5]) * 1024) + 16;
        }
    }
};

// img_warp[mask == 0] = GaussianBlur(img_warp)[mask == 0] (align.py:249-251).  `out` already holds a copy of
// the warped frame (AlWarp's second destination); the listed pixels (thin border strips) are replaced by the
// full 2-D fixed-point sum (sum_ij k_i k_j p_ij + half) >> shift -- exact integers, so the order of the taps
// is free: one WARP per pixel, the kernel columns split over its lanes, shuffle reduction.  (One thread per pixel
// left a long tail: all the heavy pixels sit in a few border CTAs.)  The host build runs item(i, 0, 1).
template <typename T>
struct AlBlurFillList {
    const T *warped; long long wpitch; int h, w; const int *list;
    int taps; int k[AL_MAX_BLUR_TAPS];
    T *out; long long opitch;
    SS_MDEV void item(int i, int lane, int nlanes) const
    {
        const int idx = list[i];
        const int y = idx / w, x = idx - y * w, r = taps / 2;
        // a lane owns whole kernel columns: 32-bit column sums sum_i k_i p_ij (< 2^32 even for uint16), one
        // 64-bit multiply by k_j per column
        unsigned long long acc0 = 0, acc1 = 0, acc2 = 0;
        const bool inner_y = y >= r && y + r < h;
        for (int tj = lane; tj < taps; tj += nlanes) {
            const int xo = 3 * dm_reflect101(x - r + tj, w);
            unsigned c0 = 0, c1 = 0, c2 = 0;
            for (int ti = 0; ti < taps; ++ti) {
                const int yy = inner_y ? y - r + ti : dm_reflect101(y - r + ti, h);
                const T *p = (const T *)((const char *)warped + (long long)yy * wpitch) + xo;
                const unsigned ki = (unsigned)k[ti];
                c0 += ki * p[0]; c1 += ki * p[1]; c2 += ki * p[2];
            }
            const unsigned long long kj = (unsigned)k[tj];
            acc0 += kj * c0; acc1 += kj * c1; acc2 += kj * c2;
        }
        acc0 = al_lane_sum(acc0); acc1 = al_lane_sum(acc1); acc2 = al_lane_sum(acc2);
        if (lane == 0) {
            const int shift = sizeof(T) == 1 ? 16 : 32;
            T *o = (T *)((char *)out + (long long)y * opitch) + 3 * x;
            o[0] = (T)((acc0 + (1ull << (shift - 1))) >> shift);
            o[1] = (T)((acc1 + (1ull << (shift - 1))) >> shift);
            o[2] = (T)((acc2 + (1ull << (shift - 1))) >> shift);
        }
    }
};

