// ss_al_engine.h -- launch sequence of the AlignFrames warp stage (reference align.py:216-251) against the same
// tiny backend as ss_dm_engine.h (GPU in the product, plain loops in tests/emu/emu_align.cpp).  The float64
// matrix inversion and the affine coordinate tables are built here on the host exactly like cv2 does;
// everything per pixel is in the functors of ss_al_kernels.cuh.
#pragma once
#include "ss_al_kernels.cuh"
#include "ss_dm_engine.h"
#include "ss_ecc_engine.h"

enum { AL_RIGID = 0, AL_HOMOGRAPHY = 1 };
enum { AL_BORDER_CONSTANT = 0, AL_BORDER_REPLICATE = 1, AL_BORDER_REPLICATE_BLUR = 2 };

template <class Backend>
struct AlEngine {
    Backend be;
    int h = 0, w = 0, bps = 0;
    size_t pitch = 0, frame_bytes = 0;
    uint8_t *d_src = nullptr, *d_next = nullptr;   // the frame being aligned / the frame uploading behind it (prefetch)
    uint8_t *d_warp = nullptr, *d_out = nullptr;   // outputs of the current warp; d_warp_b / d_out_b: of the previous one
    uint8_t *d_warp_b = nullptr, *d_out_b = nullptr;
    bool next_pending = false;
    int *d_tab = nullptr, *d_fill = nullptr;      // d_fill[0] = number of pixels to fill, d_fill[1..] = their indices
    const uint8_t *result = nullptr;
    char err[256];
    EccEstimator<Backend> ecc;          // allocated on first use
    bool ecc_ready = false, src_valid = false;

    int fail(int code, const char *msg) { snprintf(err, sizeof(err), "%s", msg); return code; }

    int create(int width, int height, int bytes_per_sample)
    {
        err[0] = 0;
        w = width; h = height; bps = bytes_per_sample;
        if (w < 1 || h < 1 || w > 32767 || h > 32767 || (bps != 1 && bps != 2))
            return fail(DM_ERR_INVALID_OPTION, "invalid frame geometry");
        pitch = (size_t)w * 3 * bps;
        frame_bytes = pitch * h;
        int rc;
        if ((rc = be.alloc((void **)&d_src, frame_bytes))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_next, frame_bytes))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_warp, frame_bytes))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_out, frame_bytes))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_warp_b, frame_bytes))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_out_b, frame_bytes))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_fill, sizeof(int) * ((size_t)w * h + 1)))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_tab, sizeof(int) * 2 * ((size_t)w + h)))) return fail(rc, be.last_error());
        return DM_OK;
    }

    void destroy()
    {
        be.sync();
        if (ecc_ready) { ecc.destroy(); ecc_ready = false; }
        void *ptrs[] = {d_src, d_next, d_warp, d_out, d_warp_b, d_out_b, d_fill, d_tab};
        for (void *p : ptrs) if (p) be.free(p);
        d_src = d_next = d_warp = d_out = d_warp_b = d_out_b = nullptr; d_tab = d_fill = nullptr;
    }

    template <typename T>
    int run(const double *m, int transform, int border_mode, const double *bval, const int *kernel, int taps)
    {
        AlWarp<T> k;
        memset(&k, 0, sizeof(k));
        k.src = (const T *)d_src; k.spitch = (long long)pitch; k.h = h; k.w = w;
        k.dst = (T *)d_warp; k.dpitch = (long long)pitch;
        const bool fill = border_mode == AL_BORDER_REPLICATE_BLUR;
        k.dst2 = fill ? (T *)d_out : nullptr;
        k.fill_count = fill ? d_fill : nullptr;
        k.fill_list = fill ? d_fill + 1 : nullptr;
        int rc;
        if (fill && (rc = be.zero(d_fill, sizeof(int)))) return rc;
        k.border_const = border_mode == AL_BORDER_CONSTANT;
        const double maxv = bps == 1 ? 255.0 : 65535.0;
        for (int c = 0; c < 3; ++c) {       // saturate_cast<T>(borderValue[c])
            double v = bval ? nearbyint(bval[c]) : 0.0;
            k.bval[c] = (int)(v < 0 ? 0 : (v > maxv ? maxv : v));
        }
        if (transform == AL_RIGID) {
            // cv2.warpAffine: invert the 2x3 matrix in float64, then fixed-point tables with AB_BITS = 10
            double M[6];
            memcpy(M, m, sizeof(M));
            double D = M[0] * M[4] - M[1] * M[3];
            D = D != 0 ? 1. / D : 0;
            const double A11 = M[4] * D, A22 = M[0] * D;
            M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
            const double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
            M[2] = b1; M[5] = b2;
            AlAffineTables t;
            for (int i = 0; i < 6; ++i) t.M[i] = M[i];
            t.w = w; t.h = h; t.ad = d_tab; t.bd = d_tab + w; t.x0 = d_tab + 2 * w; t.y0 = d_tab + 2 * w + h;
            if ((rc = be.launch(t, w > h ? w : h, 1, 1))) return rc;
            k.persp = 0;
            k.adelta = d_tab; k.bdelta = d_tab + w; k.x0 = d_tab + 2 * w; k.y0 = d_tab + 2 * w + h;
        } else {
            // cv2.warpPerspective: cv::invert of the 3x3 matrix (adjugate / determinant in float64)
            const double *S = m;
            double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) +
                       S[2] * (S[3] * S[7] - S[4] * S[6]);
            if (d == 0) return DM_ERR_INVALID_OPTION;
            d = 1. / d;
            k.m[0] = (S[4] * S[8] - S[5] * S[7]) * d; k.m[1] = (S[2] * S[7] - S[1] * S[8]) * d;
            k.m[2] = (S[1] * S[5] - S[2] * S[4]) * d; k.m[3] = (S[5] * S[6] - S[3] * S[8]) * d;
            k.m[4] = (S[0] * S[8] - S[2] * S[6]) * d; k.m[5] = (S[2] * S[3] - S[0] * S[5]) * d;
            k.m[6] = (S[3] * S[7] - S[4] * S[6]) * d; k.m[7] = (S[1] * S[6] - S[0] * S[7]) * d;
            k.m[8] = (S[0] * S[4] - S[1] * S[3]) * d;
            k.persp = 1;
            k.float_persp = bps == 2;
        }
        if ((rc = be.launch(k, w, h, 1))) return rc;
        result = d_warp;
        if (fill) {
            AlBlurFillList<T> b;
            memset(&b, 0, sizeof(b));
            b.warped = (const T *)d_warp; b.wpitch = (long long)pitch; b.h = h; b.w = w; b.list = d_fill + 1;
            b.taps = taps;
            for (int i = 0; i < taps; ++i) b.k[i] = kernel[i];
            b.out = (T *)d_out; b.opitch = (long long)pitch;
            if ((rc = be.launch_items(b, d_fill))) return rc;
            result = d_out;
        }
        return DM_OK;
    }

    int ecc_prepare()
    {
        if (ecc_ready) return DM_OK;
        const int rc = ecc.init(&be, h, w, bps);
        if (rc) { snprintf(err, sizeof(err), "%s", ecc.err); ecc.destroy(); return rc; }
        ecc_ready = true;
        return DM_OK;
    }

    // the frame every later frame is aligned to (AlignFrames.run_frame: process.img_ref(ref_idx), align.py:283-287)
    int ecc_set_reference(const void *src, size_t spitch, int on_device)
    {
        if (!src) return fail(DM_ERR_INVALID_OPTION, "null argument");
        if (spitch < pitch) return fail(DM_ERR_INVALID_OPTION, "row pitch smaller than a row");
        int rc = ecc_prepare();
        if (rc) return rc;
        if ((rc = be.upload_frame(d_src, pitch, src, spitch, pitch, h, on_device))) return fail(rc, be.last_error());
        src_valid = false;
        if ((rc = ecc.set_reference(d_src, pitch))) return fail(rc, ecc.err);
        rc = be.sync();
        return rc ? fail(rc, be.last_error()) : DM_OK;
    }

    // gradient-based replacement of detect_and_compute + find_transform (align.py:89-148) for ALIGN_RIGID; the
    // uploaded frame stays in d_src so that warp(src = NULL, ...) can warp it without a second upload
    int ecc_estimate(const void *src, size_t spitch, int on_device, int finest_level, int max_iters, double eps_px,
                     const double *init_m, double *m_out, double *info)
    {
        if ((!src && !src_valid) || !m_out) return fail(DM_ERR_INVALID_OPTION, "null argument");
        if (src && spitch < pitch) return fail(DM_ERR_INVALID_OPTION, "row pitch smaller than a row");
        int rc = ecc_prepare();
        if (rc) return rc;
        be.mark_start();
        if (src) {      // src == NULL: the frame made current by use_prefetched()
            if ((rc = be.upload_frame(d_src, pitch, src, spitch, pitch, h, on_device))) return fail(rc, be.last_error());
            src_valid = true;
        }
        if ((rc = ecc.estimate(d_src, pitch, finest_level, max_iters, eps_px, init_m, m_out, info))) return fail(rc, ecc.err);
        be.mark_end();
        return DM_OK;
    }

    // Pipelining (config 5): start the upload of the NEXT frame from pinned host memory into the second source
    // buffer; it runs on the copy stream behind whatever the compute stream is doing (the ECC fit and the warp of
    // the current frame).  use_prefetched() then makes it the current frame (ecc_estimate / warp with src = NULL).
    int prefetch(const void *src, size_t spitch)
    {
        if (!src) return fail(DM_ERR_INVALID_OPTION, "null argument");
        if (spitch < pitch) return fail(DM_ERR_INVALID_OPTION, "row pitch smaller than a row");
        if (next_pending) return fail(DM_ERR_STATE, "a prefetched frame is waiting: call ssal_use_prefetched first");
        const int rc = be.upload_frame_behind(d_next, pitch, src, spitch, pitch, h);
        if (rc) return fail(rc, be.last_error());
        next_pending = true;
        return DM_OK;
    }
    int use_prefetched()
    {
        if (!next_pending) return fail(DM_ERR_STATE, "no prefetched frame");
        uint8_t *t = d_src; d_src = d_next; d_next = t;
        next_pending = false;
        const int rc = be.wait_uploads();
        if (rc) return fail(rc, be.last_error());
        src_valid = true;
        return DM_OK;
    }

    // replaces align.py:228-251 for one frame; dst_host may be NULL (result stays on the device)
    int warp(const void *src, size_t spitch, int src_on_device, const double *m, int transform, int border_mode,
             const double *bval, const int *kernel, int taps, void *dst_host, size_t dpitch)
    {
        if (!m || (!src && !src_valid)) return fail(DM_ERR_INVALID_OPTION, "null argument");
        if (transform != AL_RIGID && transform != AL_HOMOGRAPHY) return fail(DM_ERR_INVALID_OPTION, "transform");
        if (border_mode < 0 || border_mode > 2) return fail(DM_ERR_INVALID_OPTION, "border_mode");
        if (border_mode == AL_BORDER_REPLICATE_BLUR) {
            if (!kernel || taps < 1 || taps > AL_MAX_BLUR_TAPS || !(taps & 1))
                return fail(DM_ERR_INVALID_OPTION, "blur kernel: odd number of taps, at most 63");
            long long sum = 0;
            for (int i = 0; i < taps; ++i) sum += kernel[i];
            if (sum != (bps == 1 ? 256 : 65536)) return fail(DM_ERR_INVALID_OPTION, "blur kernel must sum to 1.0 in fixed point");
        }
        if (src && spitch < pitch) return fail(DM_ERR_INVALID_OPTION, "row pitch smaller than a row");
        be.mark_start();
        int rc = 0;
        if (src) {      // src == NULL: warp the frame the last ecc_estimate uploaded
            rc = be.upload_frame(d_src, pitch, src, spitch, pitch, h, src_on_device);
            if (rc) return fail(rc, be.last_error());
            src_valid = true;
        }
        // the previous result stays intact while this warp runs (its consumer may still be copying it)
        { uint8_t *t = d_warp; d_warp = d_warp_b; d_warp_b = t; t = d_out; d_out = d_out_b; d_out_b = t; }
        rc = bps == 1 ? run<uint8_t>(m, transform, border_mode, bval, kernel, taps)
                      : run<uint16_t>(m, transform, border_mode, bval, kernel, taps);
        if (rc) return fail(rc, rc == DM_ERR_INVALID_OPTION ? "singular transform" : be.last_error());
        be.mark_end();
        if (dst_host) {
            if (dpitch < pitch) return fail(DM_ERR_INVALID_OPTION, "output row pitch smaller than a row");
            rc = be.download2d(dst_host, dpitch, result, pitch, pitch, h);
        } else rc = be.sync();
        return rc ? fail(rc, be.last_error()) : DM_OK;
    }
};
