// ss_ecc_engine.h -- host side of the ECC estimator (see ss_ecc_kernels.cuh): gray pyramids of the reference
// and the moving frame and the coarse-to-fine loop.  The iterations of a level -- per pass the float64 algebra on the
// 28 sums (zero-mean correction, 4x4 solve, ECC lambda, parameter update; ecc_update in ss_ecc_kernels.cuh) -- run on
// the device inside one cooperative launch per level.
#pragma once
#include <math.h>
#include <string.h>
#include <vector>
#include "ss_ecc_kernels.cuh"

#define ECC_MAX_LEVELS 12
#define ECC_SEARCH 4            /* coarse search: whole-pixel shifts in [-4, 4]^2 at the coarsest level */
#define ECC_SEARCH_N ((2 * ECC_SEARCH + 1) * (2 * ECC_SEARCH + 1))
/* Step size (pixels of the level) at which a level above the finest one stops: it only has to land inside the basin
 * of the next level (about a pixel there = half a pixel here).  0.02 kept the 64 x 48 level iterating to max_iters on
 * its own rounding noise (30 of the 58 iterations of a 4096 x 3072 fit, tools/pipe_once.py). */
#ifndef ECC_COARSE_EPS
#define ECC_COARSE_EPS 0.1
#endif

template <class Backend>
struct EccEstimator {
    Backend *be = nullptr;
    int h = 0, w = 0, bps = 0, L = 0;
    int lh[ECC_MAX_LEVELS], lw[ECC_MAX_LEVELS];
    float *ref[ECC_MAX_LEVELS], *mov[ECC_MAX_LEVELS];
    float *gx = nullptr, *gy = nullptr;
    double *d_sums = nullptr;
    EccState *d_state = nullptr;     // iteration state of the estimate in flight (device memory)
    unsigned int *d_bar = nullptr;   // counter of the level kernel's grid barrier
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
        if ((rc = be->alloc((void **)&d_state, sizeof(EccState)))) return fail(rc, be->last_error());
        if ((rc = be->alloc((void **)&d_bar, 64))) return fail(rc, be->last_error());
        return DM_OK;
    }

    void destroy()
    {
        if (!be) return;
        for (int l = 0; l < L; ++l) { if (ref[l]) be->free(ref[l]); if (mov[l]) be->free(mov[l]); }
        if (gx) be->free(gx);
        if (gy) be->free(gy);
        if (d_sums) be->free(d_sums);
        if (d_state) be->free(d_state);
        if (d_bar) be->free(d_bar);
        d_state = nullptr; d_bar = nullptr;
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
    // The iterations of every level run on the device (Backend::ecc_run_level: one cooperative launch per level,
    // state in device memory); the host synchronises twice per frame: once for the exhaustive coarse search, once
    // for the final state.
    int estimate(const void *d_mov, size_t pitch, int finest_level, int max_iters, double eps_px, const double *init_m,
                 double *m_out, double *info)
    {
        if (!have_ref) return fail(DM_ERR_STATE, "ecc: no reference frame set");
        if (finest_level < 0 || finest_level >= L) finest_level = L > 1 ? 1 : 0;
        if (max_iters < 1) max_iters = 30;
        if (eps_px <= 0) eps_px = 1e-3;
        int rc = build_pyramid(d_mov, pitch, mov);
        if (rc) return rc;
        EccState st;
        memset(&st, 0, sizeof(st));
        st.a = 1;
        if (init_m) {   // invert the forward similarity-like 2x3 matrix
            const double det = init_m[0] * init_m[4] - init_m[1] * init_m[3];
            if (det == 0) return fail(DM_ERR_INVALID_OPTION, "ecc: singular initial transform");
            const double i00 = init_m[4] / det, i01 = -init_m[1] / det, i10 = -init_m[3] / det, i11 = init_m[0] / det;
            st.a = 0.5 * (i00 + i11); st.b = 0.5 * (i10 - i01);
            st.dx = -(i00 * init_m[2] + i01 * init_m[5]); st.dy = -(i10 * init_m[2] + i11 * init_m[5]);
        }
        const double s0 = ldexp(1.0, -(L - 1));
        st.dx *= s0; st.dy *= s0;
        int searched = 0;
        for (int l = L - 1; l >= finest_level; --l) {
            const int hh = lh[l], ww = lw[l];
            EccGradient g{mov[l], hh, ww, gx, gy};
            if ((rc = be->launch(g, ww, hh, 1))) return fail(rc, be->last_error());
            EccLevelArgs la;
            la.ref = ref[l]; la.mov = mov[l]; la.gx = gx; la.gy = gy; la.h = hh; la.w = ww;
            la.max_iters = max_iters; la.use_state_t = 0; la.scale_up = l > finest_level ? 1 : 0;
            // coarse levels only have to bring the next level into its basin
            la.eps = l > finest_level ? (eps_px > ECC_COARSE_EPS ? eps_px : ECC_COARSE_EPS) : eps_px;
            if (l == L - 1) {
                if (!init_m) {
                    // coarsest level: exhaustive search over whole-pixel shifts (+-4 px here = +-4 * 2^(L-1) px at full
                    // resolution) so that the iterations start inside the basin of the right optimum even on
                    // content without low frequencies; all candidates are evaluated with one synchronisation
                    const double cx = 0.5 * (ww - 1), cy = 0.5 * (hh - 1);
                    const double tx = st.dx + st.a * cx - st.b * cy - cx, ty = st.dy + st.b * cx + st.a * cy - cy;
                    std::vector<double> sums((size_t)ECC_SEARCH_N * ECC_NSUMS);
                    EccParams base;
                    base.ref = ref[l]; base.mov = mov[l]; base.gx = gx; base.gy = gy; base.h = hh; base.w = ww;
                    base.a = (float)st.a; base.b = (float)st.b; base.tx = (float)tx; base.ty = (float)ty;
                    base.cx = (float)cx; base.cy = (float)cy; base.mean_r = 0.f; base.mean_w = 0.f;
                    searched = ECC_SEARCH_N;
                    if ((rc = be->ecc_search(base, tx, ty, ECC_SEARCH, d_sums, sums.data())))
                        return fail(rc, be->last_error());
                    double best = -3, btx = tx, bty = ty;
                    for (int c = 0; c < ECC_SEARCH_N; ++c) {
                        const double *Q = sums.data() + (size_t)c * ECC_NSUMS;
                        const double n = Q[27];
                        double r = -2.0;
                        if (n >= 64) {
                            const double mw = Q[25] / n, mr = Q[26] / n;
                            const double sww = Q[22] - n * mw * mw, srw = Q[23] - n * mw * mr, srr = Q[24] - n * mr * mr;
                            r = sww > 0 && srr > 0 ? srw / sqrt(sww * srr) : -2.0;
                        }
                        if (r > best) {
                            const EccParams p = ecc_search_candidate(base, tx, ty, ECC_SEARCH, c);
                            best = r; btx = p.tx; bty = p.ty;
                        }
                    }
                    st.tx = btx; st.ty = bty;
                    la.use_state_t = 1;
                }
                if ((rc = be->ecc_set_state(d_state, st))) return fail(rc, be->last_error());
            }
            if ((rc = be->ecc_run_level(la, d_state, d_sums, d_bar))) return fail(rc, be->last_error());
        }
        if ((rc = be->download(&st, d_state, sizeof(st)))) return fail(rc, be->last_error());
        if (st.fail == ECC_FAIL_NO_OVERLAP) return fail(DM_ERR_STATE, "ecc: the frames do not overlap");
        if (st.fail == ECC_FAIL_FLAT) return fail(DM_ERR_STATE, "ecc: flat image, nothing to align");
        if (st.fail) return fail(DM_ERR_STATE, "ecc: singular normal equations");
        double a = st.a, b = st.b, dx = st.dx, dy = st.dy;
        const double s = ldexp(1.0, finest_level);
        dx *= s; dy *= s;
        // forward transform = inverse of (x -> A x + d), A = [[a, -b], [b, a]]
        const double det = a * a + b * b;
        const double ia = a / det, ib = -b / det;
        m_out[0] = ia; m_out[1] = -ib; m_out[2] = -(ia * dx - ib * dy);
        m_out[3] = ib; m_out[4] = ia; m_out[5] = -(ib * dx + ia * dy);
        if (info) { info[0] = st.rho; info[1] = st.total + searched; info[2] = st.overlap; info[3] = st.last; }
        return DM_OK;
    }
};
