// ss_ecc_engine.h -- host side of the ECC estimator (see ss_ecc_kernels.cuh): gray pyramids of the reference
// and the moving frame, the coarse-to-fine loop, and per iteration the float64 algebra on the 28 sums that
// the device pass returns (zero-mean correction, 4x4 solve, ECC lambda, parameter update).
#pragma once
#include <math.h>
#include <vector>
#include "ss_ecc_kernels.cuh"

#define ECC_MAX_LEVELS 12
#define ECC_SEARCH 4            /* coarse search: whole-pixel shifts in [-4, 4]^2 at the coarsest level */
#define ECC_SEARCH_N ((2 * ECC_SEARCH + 1) * (2 * ECC_SEARCH + 1))

static inline bool ecc_solve4(const double *Hs /*10 upper-triangular*/, const double *b, double *x)
{
    double A[4][5];
    int k = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) { A[i][j] = A[j][i] = Hs[k++]; }
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

template <class Backend>
struct EccEstimator {
    Backend *be = nullptr;
    int h = 0, w = 0, bps = 0, L = 0;
    int lh[ECC_MAX_LEVELS], lw[ECC_MAX_LEVELS];
    float *ref[ECC_MAX_LEVELS], *mov[ECC_MAX_LEVELS];
    float *gx = nullptr, *gy = nullptr;
    double *d_sums = nullptr;
    bool have_ref = false;
    char err[256];

    int fail(int code, const char *msg) { snprintf(err, sizeof(err), "%s", msg); return code; }

    int init(Backend *backend, int height, int width, int bytes_per_sample)
    {
        be = backend; h = height; w = width; bps = bytes_per_sample;
        err[0] = 0;
        L = 1;
        lh[0] = h; lw[0] = w;
        while (L < ECC_MAX_LEVELS && (lh[L - 1] + 1) / 2 >= 32 && (lw[L - 1] + 1) / 2 >= 32) {
            lh[L] = (lh[L - 1] + 1) / 2; lw[L] = (lw[L - 1] + 1) / 2; ++L;
        }
        int rc;
        for (int l = 0; l < L; ++l) {
            const size_t bytes = sizeof(float) * (size_t)lh[l] * lw[l];
            if ((rc = be->alloc((void **)&ref[l], bytes))) return fail(rc, be->last_error());
            if ((rc = be->alloc((void **)&mov[l], bytes))) return fail(rc, be->last_error());
        }
        const size_t bytes0 = sizeof(float) * (size_t)h * w;
        if ((rc = be->alloc((void **)&gx, bytes0))) return fail(rc, be->last_error());
        if ((rc = be->alloc((void **)&gy, bytes0))) return fail(rc, be->last_error());
        if ((rc = be->alloc((void **)&d_sums, sizeof(double) * ECC_NSUMS * ECC_SEARCH_N))) return fail(rc, be->last_error());
        return DM_OK;
    }

    void destroy()
    {
        if (!be) return;
        for (int l = 0; l < L; ++l) { if (ref[l]) be->free(ref[l]); if (mov[l]) be->free(mov[l]); }
        if (gx) be->free(gx);
        if (gy) be->free(gy);
        if (d_sums) be->free(d_sums);
        L = 0; gx = gy = nullptr; d_sums = nullptr; be = nullptr;
    }

    // gray pyramid of a frame that sits in device memory (interleaved BGR)
    int build_pyramid(const void *d_frame, size_t pitch, float **pyr)
    {
        int rc;
        if (bps == 1) { EccGray<uint8_t> g{d_frame, (long long)pitch, h, w, pyr[0]}; rc = be->launch(g, w, h, 1); }
        else { EccGray<uint16_t> g{d_frame, (long long)pitch, h, w, pyr[0]}; rc = be->launch(g, w, h, 1); }
        for (int l = 1; !rc && l < L; ++l) {
            EccPyrDown d{pyr[l - 1], lh[l - 1], lw[l - 1], pyr[l], lh[l], lw[l]};
            rc = be->launch(d, lw[l], lh[l], 1);
        }
        return rc ? fail(rc, be->last_error()) : DM_OK;
    }

    int set_reference(const void *d_frame, size_t pitch)
    {
        const int rc = build_pyramid(d_frame, pitch, ref);
        have_ref = rc == DM_OK;
        return rc;
    }

    // Estimates the forward transform m (2x3, moving frame -> reference frame, what
    // cv2.estimateAffinePartial2D returns in the reference) of the moving frame in device memory.
    // info: [0] correlation coefficient, [1] iterations, [2] overlap fraction at the finest level, [3] last step (px)
    int estimate(const void *d_mov, size_t pitch, int finest_level, int max_iters, double eps_px, const double *init_m,
                 double *m_out, double *info)
    {
        if (!have_ref) return fail(DM_ERR_STATE, "ecc: no reference frame set");
        if (finest_level < 0 || finest_level >= L) finest_level = L > 1 ? 1 : 0;
        if (max_iters < 1) max_iters = 30;
        if (eps_px <= 0) eps_px = 1e-3;
        int rc = build_pyramid(d_mov, pitch, mov);
        if (rc) return rc;
        // parameters of the map reference -> moving in absolute level coordinates: mov = A x + d
        double a = 1, b = 0, dx = 0, dy = 0;
        if (init_m) {   // invert the forward similarity-like 2x3 matrix
            const double det = init_m[0] * init_m[4] - init_m[1] * init_m[3];
            if (det == 0) return fail(DM_ERR_INVALID_OPTION, "ecc: singular initial transform");
            const double i00 = init_m[4] / det, i01 = -init_m[1] / det, i10 = -init_m[3] / det, i11 = init_m[0] / det;
            a = 0.5 * (i00 + i11); b = 0.5 * (i10 - i01);
            dx = -(i00 * init_m[2] + i01 * init_m[5]); dy = -(i10 * init_m[2] + i11 * init_m[5]);
        }
        const double s0 = ldexp(1.0, -(L - 1));
        dx *= s0; dy *= s0;
        double rho = 0, last = 0, overlap = 0;
        int total = 0;
        for (int l = L - 1; l >= finest_level; --l) {
            const int hh = lh[l], ww = lw[l];
            EccGradient g{mov[l], hh, ww, gx, gy};
            if ((rc = be->launch(g, ww, hh, 1))) return fail(rc, be->last_error());
            const double cx = 0.5 * (ww - 1), cy = 0.5 * (hh - 1);
            // centre-relative translation: mov = A (x - c) + t + c  =>  t = d + A c - c
            double tx = dx + a * cx - b * cy - cx, ty = dy + b * cx + a * cy - cy;
            const double radius = 0.5 * sqrt((double)ww * ww + (double)hh * hh);
            double mean_r = 0, mean_w = 0;
            double S[ECC_NSUMS];
            auto params = [&](double pa, double pb, double ptx, double pty) -> EccParams {
                EccParams p;
                p.ref = ref[l]; p.mov = mov[l]; p.gx = gx; p.gy = gy; p.h = hh; p.w = ww;
                p.a = (float)pa; p.b = (float)pb; p.tx = (float)ptx; p.ty = (float)pty; p.cx = (float)cx; p.cy = (float)cy;
                p.mean_r = (float)mean_r; p.mean_w = (float)mean_w;
                return p;
            };
            auto pass = [&](double pa, double pb, double ptx, double pty) -> int {
                const EccParams p = params(pa, pb, ptx, pty);
                ++total;
                return be->ecc_accumulate_batch(&p, 1, d_sums, S);
            };
            auto correlation_of = [](const double *Q) -> double {   // zero-mean samples over the overlap; -2 = unusable
                const double n = Q[27];
                if (n < 64) return -2.0;
                const double mw = Q[25] / n, mr = Q[26] / n;
                const double sww = Q[22] - n * mw * mw, srw = Q[23] - n * mw * mr, srr = Q[24] - n * mr * mr;
                return sww > 0 && srr > 0 ? srw / sqrt(sww * srr) : -2.0;
            };
            auto correlation = [&]() -> double { return correlation_of(S); };
            if (l == L - 1 && !init_m) {
                // coarsest level: exhaustive search over whole-pixel shifts (+-4 px here = +-4 * 2^(L-1) px at full
                // resolution) so that the iterations start inside the basin of the right optimum even on
                // content without low frequencies; all candidates are evaluated with one synchronisation
                std::vector<EccParams> cand;
                std::vector<double> sums((size_t)ECC_SEARCH_N * ECC_NSUMS);
                for (int sy = -ECC_SEARCH; sy <= ECC_SEARCH; ++sy)
                    for (int sx = -ECC_SEARCH; sx <= ECC_SEARCH; ++sx) cand.push_back(params(a, b, tx + sx, ty + sy));
                total += (int)cand.size();
                if ((rc = be->ecc_accumulate_batch(cand.data(), (int)cand.size(), d_sums, sums.data())))
                    return fail(rc, be->last_error());
                double best = -3, btx = tx, bty = ty;
                for (size_t c = 0; c < cand.size(); ++c) {
                    const double r = correlation_of(sums.data() + c * ECC_NSUMS);
                    if (r > best) { best = r; btx = cand[c].tx; bty = cand[c].ty; }
                }
                tx = btx; ty = bty;
            }
            double prev_rho = -3, pa = a, pb = b, ptx = tx, pty = ty, step[4] = {0, 0, 0, 0};
            int halvings = 0;
            for (int it = 0; it < max_iters; ++it) {
                if ((rc = pass(a, b, tx, ty))) return fail(rc, be->last_error());
                const double n = S[27];
                double r_now = correlation();
                if (it > 0 && r_now < prev_rho - 1e-4 && halvings < 5) {
                    // the last step lowered the correlation (or left the frame): retreat to half of it
                    ++halvings;
                    for (int i = 0; i < 4; ++i) step[i] *= 0.5;
                    a = pa + step[0]; b = pb + step[1]; tx = ptx + step[2]; ty = pty + step[3];
                    continue;
                }
                if (n < 64) return fail(DM_ERR_STATE, "ecc: the frames do not overlap");
                halvings = 0;
                overlap = n / ((double)hh * ww);
                const double mw = S[25] / n, mr = S[26] / n;   // means of the offset values over the overlap
                double Gw[4], Gr[4];
                for (int i = 0; i < 4; ++i) { Gw[i] = S[10 + i] - mw * S[18 + i]; Gr[i] = S[14 + i] - mr * S[18 + i]; }
                const double sww = S[22] - n * mw * mw, srw = S[23] - n * mw * mr, srr = S[24] - n * mr * mr;
                if (sww <= 0 || srr <= 0) return fail(DM_ERR_STATE, "ecc: flat image, nothing to align");
                rho = srw / sqrt(sww * srr);
                prev_rho = rho; pa = a; pb = b; ptx = tx; pty = ty;
                double HiGw[4];
                if (!ecc_solve4(S, Gw, HiGw)) return fail(DM_ERR_STATE, "ecc: singular normal equations");
                double num = sww, den = srw;
                for (int i = 0; i < 4; ++i) { num -= Gw[i] * HiGw[i]; den -= Gr[i] * HiGw[i]; }
                const double lambda = den > 0 ? num / den : (srw > 0 ? sww / srw : 1.0);
                double rhs[4], dp[4];
                for (int i = 0; i < 4; ++i) rhs[i] = lambda * Gr[i] - Gw[i];
                if (!ecc_solve4(S, rhs, dp)) return fail(DM_ERR_STATE, "ecc: singular normal equations");
                last = sqrt(dp[2] * dp[2] + dp[3] * dp[3]) + radius * sqrt(dp[0] * dp[0] + dp[1] * dp[1]);
                const double cap = last > 2.0 ? 2.0 / last : 1.0;   // no pixel moves by more than ~2 px per step
                for (int i = 0; i < 4; ++i) step[i] = dp[i] * cap;
                a += step[0]; b += step[1]; tx += step[2]; ty += step[3];
                // the offsets only condition the float sums; keep them fixed after the first pass of a level
                if (it == 0) { mean_r += mr; mean_w += mw; }
                // coarse levels only have to bring the next level into its basin
                if (last < (l > finest_level ? (eps_px > 0.02 ? eps_px : 0.02) : eps_px)) break;
            }
            dx = tx - a * cx + b * cy + cx; dy = ty - b * cx - a * cy + cy;
            if (l > finest_level) { dx *= 2; dy *= 2; }
        }
        const double s = ldexp(1.0, finest_level);
        dx *= s; dy *= s;
        // forward transform = inverse of (x -> A x + d), A = [[a, -b], [b, a]]
        const double det = a * a + b * b;
        const double ia = a / det, ib = -b / det;
        m_out[0] = ia; m_out[1] = -ib; m_out[2] = -(ia * dx - ib * dy);
        m_out[3] = ib; m_out[4] = ia; m_out[5] = -(ib * dx + ia * dy);
        if (info) { info[0] = rho; info[1] = total; info[2] = overlap; info[3] = last; }
        return DM_OK;
    }
};
