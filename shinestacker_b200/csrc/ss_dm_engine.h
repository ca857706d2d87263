// ss_dm_engine.h -- launch sequence of the depth-map stacker (reference depth_map.py:64-123), written
// against a tiny backend (allocate / copy / launch a pixel functor) so that the very same sequence runs
// on the GPU in the product (CudaBackend, ss_dm_lib.cuh) and in plain loops in the CPU emulation used by
// the tests (tests/emu/emu_depthmap.cpp).  No arithmetic lives here: every pixel is computed by the
// functors of ss_dm_kernels.cuh.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "ss_dm_kernels.cuh"

#define DM_BLEND_BATCH_MAX 8
enum { DM_OK = 0, DM_ERR_INVALID_OPTION = -1, DM_ERR_CUDA = -4, DM_ERR_STATE = -6, DM_ERR_NOMEM = -7 };

struct DmOptions {          // DepthMapStack.__init__ (depth_map.py:11-26) + frame geometry
    int width, height, bytes_per_sample;
    int map_type;           // 0 'average', 1 'max'
    int energy;             // 0 'laplacian', 1 'sobel'
    int kernel_size, blur_size, smooth_size;
    double temperature;
    int levels;
    int max_frames;
};

// cv2.getGaussianKernel(n, sigma <= 0, CV_32F): fixed tables up to 9 taps, else sigma = 0.3*((n-1)*0.5-1)+0.8
static inline void dm_gaussian_kernel(int n, std::vector<float> &k)
{
    static const float t1[] = {1.f}, t3[] = {0.25f, 0.5f, 0.25f}, t5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f},
                       t7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f},
                       t9[] = {4 / 256.f, 13 / 256.f, 30 / 256.f, 51 / 256.f, 60 / 256.f, 51 / 256.f, 30 / 256.f,
                               13 / 256.f, 4 / 256.f};
    const float *tabs[] = {t1, t3, t5, t7, t9};
    k.resize(n);
    if (n <= 9) { memcpy(k.data(), tabs[n / 2], n * sizeof(float)); return; }
    const double sigma = 0.3 * ((n - 1) * 0.5 - 1) + 0.8, s2 = -0.5 / (sigma * sigma);
    std::vector<double> d(n);
    double sum = 0;
    for (int i = 0; i < n; ++i) { const double x = i - (n - 1) * 0.5; d[i] = exp(s2 * x * x); sum += d[i]; }
    for (int i = 0; i < n; ++i) k[i] = (float)(d[i] / sum);
}

// cv2.getDerivKernels / getSobelKernels(order, ksize > 3): integer binomial construction
static inline void dm_sobel_kernel(int ksize, int order, std::vector<float> &k)
{
    std::vector<long long> ker(ksize + 1, 0);
    ker[0] = 1;
    for (int i = 0; i < ksize - order - 1; ++i) {
        long long oldval = ker[0];
        for (int j = 1; j <= ksize; ++j) { const long long nv = ker[j] + ker[j - 1]; ker[j - 1] = oldval; oldval = nv; }
    }
    for (int i = 0; i < order; ++i) {
        long long oldval = -ker[0];
        for (int j = 1; j <= ksize; ++j) { const long long nv = ker[j - 1] - ker[j]; ker[j - 1] = oldval; oldval = nv; }
    }
    k.resize(ksize);
    for (int i = 0; i < ksize; ++i) k[i] = (float)ker[i];
}

static inline DmTaps dm_make_taps(const std::vector<float> &k)
{
    DmTaps t;
    memset(&t, 0, sizeof(t));
    const int n = (int)k.size();
    t.r = n / 2;
    t.anti = (t.r > 0 && k[t.r] == 0.f && k[0] == -k[n - 1] && k[0] != 0.f) ? 1 : 0;
    for (int j = 0; j <= t.r; ++j) t.k[j] = k[t.r + j];
    return t;
}

template <class Backend>
struct DmEngine {
    Backend be;
    DmOptions o;
    int h = 0, w = 0, bps = 0, L = 1, n = 0, finished = 0, blended = 0, collapsed = 0;
    long long plane = 0;
    size_t frame_pitch = 0, frame_bytes = 0;
    char err[256];
    uint8_t *d_frames = nullptr, *d_out = nullptr;
    float *d_energy = nullptr, *d_smooth = nullptr, *d_t1 = nullptr, *d_t2 = nullptr, *d_t3 = nullptr, *d_sw = nullptr;
    float *weights = nullptr;      // the buffer that holds the focus maps after finish_begin
    unsigned *d_stats = nullptr;        // per frame (min, max) bits of the raw energy, folded by finish_begin
    unsigned *d_slots = nullptr;        // per frame DM_STAT_SLOTS (min, max) pairs the energy kernel updates
    std::vector<int> lh, lw;
    std::vector<float *> gp, acc;  // gp[l] (l >= 1): 4 planes; acc[l]: 3 planes
    std::vector<float> h_sw;       // bilateral space weights, (2r+1)^2, 0 outside the disc
    DmTaps t_blur, t_a, t_b;
    int bil_r = 0;
    float gmax = 0.f;
    int batch = DM_BLEND_BATCH_MAX, launched_blend = 0;   // frames per blend launch (SSDM_BLEND_BATCH=1..8)
    int fold_norm = 0, moments = 3, tiled_pyrdown = 1;
    // frames of one stack sharded over ranks (shard_* below): this rank holds a contiguous run of the frames
    int world = 1, rank = 0, root = 0, stats_folded = 0;
    float local_max = 0.f;
    float *d_psum = nullptr, *d_pmax = nullptr;
    int rel_done = 0;
    const float *peer_psum[DM_MAX_RANKS] = {};
    const float *peer_pmax[DM_MAX_RANKS] = {};
    const float *peer_acc[DM_MAX_RANKS][16] = {};
    int fused_energy = 1;          // SSDM_ENERGY_FUSED=0: the four streaming passes instead of the tiled kernel

    int fail(int code, const char *msg) { snprintf(err, sizeof(err), "%s", msg); return code; }

    int create(const DmOptions &opt)
    {
        o = opt;
        err[0] = 0;
        h = o.height; w = o.width; bps = o.bytes_per_sample;
        if (w < 1 || h < 1 || (bps != 1 && bps != 2)) return fail(DM_ERR_INVALID_OPTION, "invalid frame geometry");
        if (o.map_type != 0 && o.map_type != 1) return fail(DM_ERR_INVALID_OPTION, "map_type: valid values are average and max");
        if (o.energy != 0 && o.energy != 1) return fail(DM_ERR_INVALID_OPTION, "energy: valid values are laplacian and sobel");
        if (o.levels < 1 || o.levels > 16) return fail(DM_ERR_INVALID_OPTION, "levels must be in 1..16");
        if (o.max_frames < 1 || o.max_frames > 65535) return fail(DM_ERR_INVALID_OPTION, "max_frames must be in 1..65535");
        std::vector<float> k;
        if (o.energy == 0) {
            if (o.blur_size < 1 || o.blur_size > 31 || !(o.blur_size & 1))
                return fail(DM_ERR_INVALID_OPTION, "blur_size must be odd and in 1..31");
            // cv2.Laplacian(float32 -> CV_64F) with ksize 1/3 raises in the reference's cv2; ksize must be odd
            if (o.kernel_size < 5 || o.kernel_size > 31 || !(o.kernel_size & 1))
                return fail(DM_ERR_INVALID_OPTION, "kernel_size must be odd and in 5..31 (cv2.Laplacian float32 -> CV_64F)");
            dm_gaussian_kernel(o.blur_size, k); t_blur = dm_make_taps(k);
            dm_sobel_kernel(o.kernel_size, 2, k); t_a = dm_make_taps(k);
            dm_sobel_kernel(o.kernel_size, 0, k); t_b = dm_make_taps(k);
        } else {
            k.assign(1, 1.f); t_blur = dm_make_taps(k);
            const float kd[] = {-1.f, 0.f, 1.f}, ks[] = {1.f, 2.f, 1.f};
            k.assign(kd, kd + 3); t_a = dm_make_taps(k);
            k.assign(ks, ks + 3); t_b = dm_make_taps(k);
        }
        const char *bb = getenv("SSDM_BLEND_BATCH");
        batch = bb ? atoi(bb) : DM_BLEND_BATCH_MAX;
        if (batch < 1) batch = 1;
        if (batch > DM_BLEND_BATCH_MAX) batch = DM_BLEND_BATCH_MAX;
        if (batch > o.max_frames) batch = o.max_frames;
        const char *pt = getenv("SSDM_PYRDOWN_TILED");
        tiled_pyrdown = !(pt && pt[0] == '0');
        const char *fe = getenv("SSDM_ENERGY_FUSED");
        fused_energy = !(fe && fe[0] == '0');
        L = o.levels;
        lh.assign(L, 0); lw.assign(L, 0);
        lh[0] = h; lw[0] = w;
        for (int l = 1; l < L; ++l) { lh[l] = (lh[l - 1] + 1) / 2; lw[l] = (lw[l - 1] + 1) / 2; }
        plane = (long long)h * w;
        frame_pitch = (size_t)w * 3 * bps;
        frame_bytes = frame_pitch * h;
        const size_t nf = (size_t)o.max_frames;
        int rc;
        if ((rc = be.alloc((void **)&d_frames, nf * frame_bytes))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_out, frame_bytes))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_energy, nf * plane * sizeof(float)))) return fail(rc, be.last_error());
        if (o.smooth_size > 0 && (rc = be.alloc((void **)&d_smooth, nf * plane * sizeof(float)))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_t1, plane * sizeof(float)))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_t2, plane * sizeof(float)))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_t3, plane * sizeof(float)))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_stats, nf * 2 * sizeof(unsigned)))) return fail(rc, be.last_error());
        if ((rc = be.alloc((void **)&d_slots, nf * 2 * DM_STAT_SLOTS * sizeof(unsigned)))) return fail(rc, be.last_error());
        gp.assign(L, nullptr); acc.assign(L, nullptr);
        for (int l = 0; l < L; ++l) {
            const size_t px = (size_t)lh[l] * lw[l];
            if (l > 0 && (rc = be.alloc((void **)&gp[l], (size_t)batch * 4 * px * sizeof(float)))) return fail(rc, be.last_error());
            if ((rc = be.alloc((void **)&acc[l], 3 * px * sizeof(float)))) return fail(rc, be.last_error());
        }
        if (o.smooth_size > 0) {
            // cv2.bilateralFilter: radius = max(d/2, 1); space weight exp(-r^2 / (2*25^2)) inside the disc
            bil_r = o.smooth_size / 2 > 1 ? o.smooth_size / 2 : 1;
            if (bil_r > 32) return fail(DM_ERR_INVALID_OPTION, "smooth_size must be <= 65");
            const int d = 2 * bil_r + 1;
            std::vector<float> &sw = h_sw;
            sw.assign((size_t)d * d, 0.f);
            const double gs = -0.5 / (25.0 * 25.0);
            for (int i = -bil_r; i <= bil_r; ++i)
                for (int j = -bil_r; j <= bil_r; ++j) {
                    const double r = sqrt((double)i * i + (double)j * j);
                    if (r > bil_r) continue;
                    sw[(size_t)(i + bil_r) * d + (j + bil_r)] = (float)exp(r * r * gs);
                }
            if ((rc = be.alloc((void **)&d_sw, sw.size() * sizeof(float)))) return fail(rc, be.last_error());
            if ((rc = be.upload_small(d_sw, sw.data(), sw.size() * sizeof(float)))) return fail(rc, be.last_error());
        }
        return reset();
    }

    void destroy()
    {
        be.sync();
        void *ptrs[] = {d_frames, d_out, d_energy, d_smooth, d_t1, d_t2, d_t3, d_stats, d_slots, d_sw, d_psum, d_pmax};
        for (void *p : ptrs) if (p) be.free(p);
        for (float *p : gp) if (p) be.free(p);
        for (float *p : acc) if (p) be.free(p);
        d_frames = d_out = nullptr; d_energy = d_smooth = d_t1 = d_t2 = d_t3 = d_sw = d_psum = d_pmax = nullptr; d_stats = d_slots = nullptr;
        gp.clear(); acc.clear();
    }

    int reset()
    {
        int rc = be.sync();
        if (rc) return fail(rc, be.last_error());
        n = 0; finished = 0; blended = 0; launched_blend = 0; collapsed = 0; weights = nullptr; gmax = 0.f;
        stats_folded = 0; local_max = 0.f; rel_done = 0;
        be.reset_counters();
        std::vector<unsigned> init(2 * (size_t)o.max_frames * DM_STAT_SLOTS);
        for (size_t i = 0; i < init.size(); i += 2) { init[i] = 0x7f800000u; init[i + 1] = 0u; }  // min = +inf, max = 0
        if ((rc = be.upload_small(d_slots, init.data(), init.size() * sizeof(unsigned)))) return fail(rc, be.last_error());
        return DM_OK;
    }

    template <typename InT, int RB, int RD>
    int energy_tiles(const void *frame, float *e, unsigned *st)
    {
        DmEnergyTile<InT, RB, RD> k;
        k.frame = frame; k.pitch = (long long)frame_pitch; k.h = h; k.w = w;
        k.tblur = t_blur; k.ta = t_a; k.tb = t_b; k.sobel = o.energy; k.out = e; k.stats = st;
        const int halo = t_blur.r + (t_a.r > t_b.r ? t_a.r : t_b.r);
        k.pw = DM_ET_X + 2 * ((halo + 3) / 4 * 4); k.ph = DM_ET_Y + 2 * halo;
        return be.launch_tiles(k, (w + DM_ET_X - 1) / DM_ET_X, (h + DM_ET_Y - 1) / DM_ET_Y, 1,
                               sizeof(float) * DM_ET_FLOATS((size_t)k.pw, k.ph));
    }

    template <int KMAX>
    int bilateral_tiles(float c1, float c2)
    {
        DmBilateralTile<KMAX> k;
        k.in = d_energy; k.h = h; k.w = w; k.stats = d_stats; k.gmax = gmax; k.out = d_smooth;
        double m0 = 0;
        for (int d = 0; d < 8; ++d) k.g[d] = (float)exp(d * d * (-0.5 / (25.0 * 25.0)));
        for (int i = -7; i <= 7; ++i)
            for (int j = -7; j <= 7; ++j)
                if (i * i + j * j <= 49) m0 += (double)k.g[abs(i)] * (double)k.g[abs(j)];
        k.c1 = c1; k.c2 = c2; k.m0 = (float)m0; k.norm = gmax > 0.f ? gmax : 0.f;
        return be.launch_tiles(k, (w + DM_BT_X - 1) / DM_BT_X, (h + DM_BT_Y - 1) / DM_BT_Y, n, sizeof(float) * DM_BT_FLOATS);
    }

    // pass 1 of one frame that already sits in slot i of the frame store
    template <typename InT>
    int energy_of_frame(int i)
    {
        int rc;
        const void *frame = d_frames + (size_t)i * frame_bytes;
        float *e = d_energy + (size_t)i * plane;
        unsigned *st = d_slots + 2 * DM_STAT_SLOTS * (size_t)i;
        if (fused_energy) {     // one tiled kernel (default); the four-pass form below is its cross-check
            const int rb = t_blur.r, rd = t_a.r > t_b.r ? t_a.r : t_b.r;
            if (rb == 2 && rd == 2) return energy_tiles<InT, 2, 2>(frame, e, st);
            if (rb == 0 && rd == 1) return energy_tiles<InT, 0, 1>(frame, e, st);
            return energy_tiles<InT, -1, -1>(frame, e, st);
        }
        if (o.energy == 0) {
            DmGrayRowFilter<InT> a{frame, (long long)frame_pitch, h, w, t_blur, d_t1};
            if ((rc = be.launch(a, w, h, 1))) return rc;
            DmColFilter b{d_t1, h, w, t_blur, d_t3};
            if ((rc = be.launch(b, w, h, 1))) return rc;
            DmRowPair c{d_t3, h, w, t_a, t_b, d_t1, d_t2};
            if ((rc = be.launch(c, w, h, 1))) return rc;
        } else {
            DmGrayRowFilter<InT> a{frame, (long long)frame_pitch, h, w, t_blur, d_t3};
            if ((rc = be.launch(a, w, h, 1))) return rc;
            DmRowPair c{d_t3, h, w, t_a, t_b, d_t1, d_t2};
            if ((rc = be.launch(c, w, h, 1))) return rc;
        }
        DmColCombine d{d_t1, d_t2, h, w, t_a, t_b, o.energy, e, st};
        return be.launch(d, w, h, 1);
    }

    // replaces one iteration of the first loop of focus_stack (depth_map.py:68-77) + that frame's share
    // of get_sobel_map / get_laplacian_map (depth_map.py:28-42)
    int push_frame(const void *bgr, size_t pitch, int on_device)
    {
        if (finished) return fail(DM_ERR_STATE, "push after finish: call reset first");
        if (n >= o.max_frames) return fail(DM_ERR_STATE, "more frames than max_frames");
        if (pitch < frame_pitch) return fail(DM_ERR_INVALID_OPTION, "row pitch smaller than a row");
        if (n == 0) be.mark_start();
        int rc = be.upload_frame(d_frames + (size_t)n * frame_bytes, frame_pitch, bgr, pitch, frame_pitch, h, on_device);
        if (rc) return fail(rc, be.last_error());
        rc = bps == 1 ? energy_of_frame<uint8_t>(n) : energy_of_frame<uint16_t>(n);
        if (rc) return fail(rc, be.last_error());
        ++n;
        return DM_OK;
    }

    // per-frame (min, max) of the raw energies out of the slots the energy kernels updated; the largest is local_max
    int fold_stats()
    {
        if (n == 0) return fail(DM_ERR_STATE, "no frames");
        if (finished) return fail(DM_ERR_STATE, "already finished");
        std::vector<unsigned> slots(2 * (size_t)n * DM_STAT_SLOTS), st(2 * (size_t)n);
        int rc = be.download(slots.data(), d_slots, slots.size() * sizeof(unsigned));
        if (rc) return fail(rc, be.last_error());
        unsigned mx = 0;
        for (int i = 0; i < n; ++i) {       // fold the slots of every frame; non-negative floats order like their bits
            unsigned lo = 0x7f800000u, hi = 0u;
            for (int s = 0; s < DM_STAT_SLOTS; ++s) {
                const unsigned a = slots[2 * ((size_t)i * DM_STAT_SLOTS + s)], b = slots[2 * ((size_t)i * DM_STAT_SLOTS + s) + 1];
                lo = a < lo ? a : lo; hi = b > hi ? b : hi;
            }
            st[2 * i] = lo; st[2 * i + 1] = hi;
            mx = hi > mx ? hi : mx;
        }
        if ((rc = be.upload_small(d_stats, st.data(), st.size() * sizeof(unsigned)))) return fail(rc, be.last_error());
        memcpy(&local_max, &mx, 4);
        stats_folded = 1;
        return DM_OK;
    }

    // depth_map.py:88-92: division by the maximum over the whole stack, bilateral smoothing (frames of this handle)
    int smooth(float global_max)
    {
        int rc;
        gmax = global_max;
        const char *pair_env = getenv("SSDM_BILATERAL_PAIR");     // A/B knob: "1" / "0" = the direct kernels
        const char *mom_env = getenv("SSDM_BILATERAL_MOMENTS");
        moments = mom_env && atoi(mom_env) == 5 ? 5 : 3;
        // the tiled bilateral divides by the global maximum as it loads: no normalisation pass, d_energy stays raw
        fold_norm = o.smooth_size > 0 && bil_r == 7 && !pair_env;
        if (gmax > 0.f && !fold_norm) {
            DmNormalize k{d_energy, h, w, gmax};
            if ((rc = be.launch(k, w, h, n))) return fail(rc, be.last_error());
        }
        weights = d_energy;
        if (o.smooth_size > 0) {
            const float gc = (float)(-0.5 / (25.0 * 25.0)), c1 = gc, c2 = 0.5f * gc * gc;
            if (bil_r == 7 && pair_env && pair_env[0] == '0') {
                DmBilateralSingle<7> k;
                k.in = d_energy; k.h = h; k.w = w; k.stats = d_stats; k.gmax = gmax; k.out = d_smooth;
                for (int i = 0; i < 225; ++i) { k.s0[i] = h_sw[i]; k.s1[i] = h_sw[i] * c1; k.s2[i] = h_sw[i] * c2; }
                if ((rc = be.launch(k, w, h, n))) return fail(rc, be.last_error());
            } else if (fold_norm) {
                // default: tiled kernel in moment form (SSDM_BILATERAL_PAIR=1 / 0 select the direct 149-tap kernels),
                // three moments (SSDM_BILATERAL_MOMENTS=5: the full colour polynomial)
                rc = moments == 5 ? bilateral_tiles<5>(c1, c2) : bilateral_tiles<3>(c1, 0.f);
                if (rc) return fail(rc, be.last_error());
            } else if (bil_r == 7) {
                DmBilateralFixed<7> k;
                k.in = d_energy; k.h = h; k.w = w; k.stats = d_stats; k.gmax = gmax; k.out = d_smooth;
                for (int i = 0; i < 225; ++i) { k.s0[i] = h_sw[i]; k.s1[i] = h_sw[i] * c1; k.s2[i] = h_sw[i] * c2; }
                if ((rc = be.launch(k, (w + 1) / 2, h, n))) return fail(rc, be.last_error());   // two pixels per call
            } else {
                DmBilateral k{d_energy, h, w, bil_r, d_sw, c1, c2, d_stats, gmax, d_smooth};
                if ((rc = be.launch(k, w, h, n))) return fail(rc, be.last_error());
            }
            weights = d_smooth;
        }
        return DM_OK;
    }

    // depth_map.py:88-93: global normalisation, bilateral smoothing, focus map
    int finish_begin()
    {
        int rc = fold_stats();
        if (rc) return rc;
        if ((rc = smooth(local_max))) return rc;
        DmFocusMap f{weights, h, w, n, o.map_type, (float)o.temperature};
        if ((rc = be.launch(f, w, h, 1))) return fail(rc, be.last_error());
        finished = 1;
        blended = 0; launched_blend = 0;
        return DM_OK;
    }

    // ---- frames of one stack sharded over ranks ------------------------------------------------------------------
    // Every rank pushes its own contiguous run of the frames.  Three values cross the ranks: the maximum energy (a
    // float, through the host), the per-pixel sum of the smoothed energies (each rank's partial sum, read from the
    // peers' memory) and the blended pyramid (every rank sums one band of rows out of the peers' memory, the root then
    // collects the bands).  The sums are taken
    // rank by rank, i.e. in stack order but associated differently from one long chain: float32 rounding, covered by
    // the 1-LSB contract of this stacker.  map_type 'max' (softmax) needs one more plane across the ranks: the per-pixel
    // maximum (exact in any order), then the sum of the exponentials.
    int shard_local_max(float *out)
    {
        const int rc = fold_stats();
        if (rc) return rc;
        *out = local_max;
        return DM_OK;
    }
    int shard_alloc()
    {
        int rc;
        if (!d_psum && (rc = be.alloc((void **)&d_psum, plane * sizeof(float)))) return fail(rc, be.last_error());
        if (!d_pmax && (rc = be.alloc((void **)&d_pmax, plane * sizeof(float)))) return fail(rc, be.last_error());
        return DM_OK;
    }
    int shard_finish_local(float global_max)
    {
        if (!stats_folded) return fail(DM_ERR_STATE, "shard_local_max first");
        int rc = smooth(global_max);
        if (rc) return rc;
        if ((rc = shard_alloc())) return rc;
        if (o.map_type == 0) {
            DmPartialSum k{weights, h, w, n, d_psum};
            rc = be.launch(k, w, h, 1);
        } else {
            DmPartialMax k{weights, h, w, n, d_pmax};
            rc = be.launch(k, w, h, 1);
        }
        return rc ? fail(rc, be.last_error()) : DM_OK;
    }
    // map_type 'max' only, after every rank has finished shard_finish_local (barrier): exp((e - max) / temperature)
    // of the own frames and their partial sum; another barrier, then shard_finish_weights
    int shard_softmax_rel()
    {
        if (!weights || finished) return fail(DM_ERR_STATE, "shard_finish_local first");
        if (o.map_type != 1) return DM_OK;
        DmSoftmaxRel k;
        k.e = weights; k.h = h; k.w = w; k.n = n; k.world = world; k.temperature = (float)o.temperature; k.psum = d_psum;
        for (int r = 0; r < DM_MAX_RANKS; ++r) k.pmax[r] = r < world ? peer_pmax[r] : nullptr;
        const int rc = be.launch(k, w, h, 1);
        if (rc) return fail(rc, be.last_error());
        rel_done = 1;
        return DM_OK;
    }
    // exchangeable buffers of this rank: ptrs[0] = partial sum plane, ptrs[1] = partial maximum plane,
    // ptrs[2 + l] = blended level l; count = 2 + L
    int shard_ptrs(void **ptrs)
    {
        const int rc = shard_alloc();
        if (rc) return rc;
        ptrs[0] = d_psum; ptrs[1] = d_pmax;
        for (int l = 0; l < L; ++l) ptrs[2 + l] = acc[l];
        return DM_OK;
    }
    // all[r * (2 + L) + k]: buffer k of rank r, mapped into this process (the own entries may be anything)
    int shard_connect(void *const *all, int world_, int rank_, int root_)
    {
        if (world_ < 1 || world_ > DM_MAX_RANKS || rank_ < 0 || rank_ >= world_ || root_ < 0 || root_ >= world_)
            return fail(DM_ERR_INVALID_OPTION, "world / rank / root out of range");
        void *own[18];
        const int rc = shard_ptrs(own);
        if (rc) return rc;
        world = world_; rank = rank_; root = root_;
        for (int r = 0; r < world; ++r)
            for (int k = 0; k < 2 + L; ++k) {
                const float *p = (const float *)(r == rank ? own[k] : all[(size_t)r * (2 + L) + k]);
                if (k == 0) peer_psum[r] = p; else if (k == 1) peer_pmax[r] = p; else peer_acc[r][k - 2] = p;
            }
        return DM_OK;
    }
    // after every rank has finished shard_finish_local (barrier): the focus maps of the own frames
    int shard_finish_weights()
    {
        if (!weights || finished) return fail(DM_ERR_STATE, "shard_finish_local first");
        if (o.map_type == 1 && !rel_done) return fail(DM_ERR_STATE, "map_type max: shard_softmax_rel first");
        DmFocusMapShared k;
        k.e = weights; k.h = h; k.w = w; k.n = n; k.world = world; k.softmax = o.map_type;
        for (int r = 0; r < DM_MAX_RANKS; ++r) k.psum[r] = r < world ? peer_psum[r] : nullptr;
        const int rc = be.launch(k, w, h, 1);
        if (rc) return fail(rc, be.last_error());
        finished = 1;
        blended = 0; launched_blend = 0;
        return DM_OK;
    }
    // after every rank has blended its frames (barrier): every rank reduces its band of rows of every level
    int shard_reduce_band()
    {
        if (!finished || blended != n) return fail(DM_ERR_STATE, "reduce before every frame was blended");
        if (world == 1) return DM_OK;
        for (int l = 0; l < L; ++l) {
            const int y0 = (int)((long long)rank * lh[l] / world), y1 = (int)((long long)(rank + 1) * lh[l] / world);
            if (y1 <= y0) continue;
            DmPeerBandSum k;
            for (int r = 0; r < DM_MAX_RANKS; ++r) k.acc[r] = r < world ? peer_acc[r][l] : nullptr;
            k.world = world; k.out = acc[l]; k.h = lh[l]; k.w = lw[l]; k.y0 = y0; k.y1 = y1;
            const int rc = be.launch(k, lw[l], 3 * (y1 - y0), 1);
            if (rc) return fail(rc, be.last_error());
        }
        return DM_OK;
    }
    // after every rank has reduced its band (barrier): the root collects the bands, collapses and stores; others return
    int shard_finish_end(void *out_bgr, size_t out_pitch)
    {
        if (!finished || blended != n) return fail(DM_ERR_STATE, "finish_end before every frame was blended");
        if (rank != root) { const int rc = be.sync(); return rc ? fail(rc, be.last_error()) : DM_OK; }
        if (!collapsed && world > 1) {
            for (int l = 0; l < L; ++l) {
                DmPeerBandGather k;
                for (int r = 0; r < DM_MAX_RANKS; ++r) k.acc[r] = r < world ? peer_acc[r][l] : nullptr;
                k.world = world; k.root = root; k.out = acc[l]; k.h = lh[l]; k.w = lw[l];
                const int rc = be.launch(k, lw[l], 3 * lh[l], 1);
                if (rc) return fail(rc, be.last_error());
            }
        }
        return finish_end(out_bgr, out_pitch);
    }

    // pass 2 of the nb frames i0 .. i0 + nb - 1: Gaussian pyramids of the batch (one launch per level, z = frame),
    // then one blend launch per level that adds the nb weighted Laplacians in stack order
    template <typename InT>
    int blend_batch_t(int i0, int nb)
    {
        int rc;
        DmLevelSrc<InT> s0{d_frames + (size_t)i0 * frame_bytes, (long long)frame_pitch, nullptr,
                           weights + (size_t)i0 * plane, h, w, (long long)frame_bytes, 0, plane};
        const int first = i0 == 0;
        auto px = [&](int l) { return (long long)lh[l] * lw[l]; };
        auto level = [&](int l) {
            return DmLevelSrc<float>{nullptr, 0, gp[l], gp[l] + 3 * px(l), lh[l], lw[l], 0, 4 * px(l), 4 * px(l)};
        };
        auto tiles = [](int v, int t) { return (v + t - 1) / t; };
        if (L > 1 && tiled_pyrdown) {
            DmPyrDownTile<InT> d0{s0, lh[1], lw[1], gp[1], 4 * px(1)};
            if ((rc = be.launch_tiles(d0, tiles(lw[1], DM_PD_X), tiles(lh[1], DM_PD_Y), nb, sizeof(float) * DM_PD_FLOATS))) return rc;
        } else if (L > 1) {
            DmPyrDown<InT> d0{s0, lh[1], lw[1], gp[1], 4 * px(1)};
            if ((rc = be.launch(d0, lw[1], lh[1], nb))) return rc;
        }
        for (int l = 1; l + 1 < L; ++l) {
            if (tiled_pyrdown) {
                DmPyrDownTile<float> d{level(l), lh[l + 1], lw[l + 1], gp[l + 1], 4 * px(l + 1)};
                if ((rc = be.launch_tiles(d, tiles(lw[l + 1], DM_PD_X), tiles(lh[l + 1], DM_PD_Y), nb, sizeof(float) * DM_PD_FLOATS))) return rc;
            } else {
                DmPyrDown<float> d{level(l), lh[l + 1], lw[l + 1], gp[l + 1], 4 * px(l + 1)};
                if ((rc = be.launch(d, lw[l + 1], lh[l + 1], nb))) return rc;
            }
        }
        {
            DmBlendLevel<InT> b{s0, L > 1 ? gp[1] : nullptr, L > 1 ? lh[1] : 0, L > 1 ? lw[1] : 0, L > 1 ? 4 * px(1) : 0,
                                acc[0], first, nb};
            if ((rc = be.launch(b, (w + 1) / 2, (h + 1) / 2, 1))) return rc;
        }
        for (int l = 1; l < L; ++l) {
            const bool top = l == L - 1;
            DmBlendLevel<float> b{level(l), top ? nullptr : gp[l + 1], top ? 0 : lh[l + 1], top ? 0 : lw[l + 1],
                                  top ? 0 : 4 * px(l + 1), acc[l], first, nb};
            if ((rc = be.launch(b, (lw[l] + 1) / 2, (lh[l] + 1) / 2, 1))) return rc;
        }
        return DM_OK;
    }

    // replaces one iteration of the second loop of focus_stack (depth_map.py:94-115); frames in stack order.
    // Frames are collected into batches of `batch` (the last one may be shorter) and launched together.
    int blend_frame(int i)
    {
        if (!finished) return fail(DM_ERR_STATE, "blend before finish_begin");
        if (i != blended || i >= n) return fail(DM_ERR_STATE, "frames must be blended once, in stack order");
        ++blended;
        if (blended - launched_blend < batch && blended < n) return DM_OK;
        const int i0 = launched_blend, nb = blended - launched_blend;
        const int rc = bps == 1 ? blend_batch_t<uint8_t>(i0, nb) : blend_batch_t<uint16_t>(i0, nb);
        if (rc) return fail(rc, be.last_error());
        launched_blend = blended;
        return DM_OK;
    }

    // depth_map.py:116-123: collapse the blended pyramid, clip(|x|), truncate; out_bgr is host memory
    int finish_end(void *out_bgr, size_t out_pitch)
    {
        if (!finished || blended != n) return fail(DM_ERR_STATE, "finish_end before every frame was blended");
        int rc = 0;
        if (!collapsed) {   // a second call only downloads the image again
            for (int l = L - 2; l >= 0; --l) {
                DmCollapse c{acc[l + 1], lh[l + 1], lw[l + 1], acc[l], lh[l], lw[l]};
                if ((rc = be.launch(c, lw[l], lh[l], 1))) return fail(rc, be.last_error());
            }
            if (bps == 1) {
                DmStoreOutput<uint8_t> s{acc[0], h, w, 255.f, d_out, (long long)frame_pitch};
                rc = be.launch(s, w, h, 1);
            } else {
                DmStoreOutput<uint16_t> s{acc[0], h, w, 65535.f, d_out, (long long)frame_pitch};
                rc = be.launch(s, w, h, 1);
            }
            if (rc) return fail(rc, be.last_error());
            collapsed = 1;
            be.mark_end();
        }
        if (out_bgr) {
            if (out_pitch < frame_pitch) return fail(DM_ERR_INVALID_OPTION, "output row pitch smaller than a row");
            if ((rc = be.download2d(out_bgr, out_pitch, d_out, frame_pitch, frame_pitch, h))) return fail(rc, be.last_error());
        } else if ((rc = be.sync())) return fail(rc, be.last_error());
        return DM_OK;
    }

    // whole chain again over the frames still resident in HBM (bench: "inputs already in HBM")
    int rerun_resident()
    {
        const int frames = n;
        if (frames == 0) return fail(DM_ERR_STATE, "no frames");
        int rc = reset();
        if (rc) return rc;
        be.mark_start();
        for (int i = 0; i < frames; ++i) {
            rc = bps == 1 ? energy_of_frame<uint8_t>(i) : energy_of_frame<uint16_t>(i);
            if (rc) return fail(rc, be.last_error());
            ++n;
        }
        if ((rc = finish_begin())) return rc;
        for (int i = 0; i < frames; ++i) if ((rc = blend_frame(i))) return rc;
        return finish_end(nullptr, 0);
    }

    // test hook: what = 0 normalised energy, 1 focus map (after finish_begin), 2 raw energy (before it)
    int get_plane(int what, int frame, float *out)
    {
        if (frame < 0 || frame >= n) return fail(DM_ERR_INVALID_OPTION, "frame out of range");
        const float *src = nullptr;
        if (what == 2 && !finished) src = d_energy;
        else if (what == 0 && finished && weights != d_energy) src = d_energy;
        else if (what == 1 && finished) src = weights;
        if (!src) return fail(DM_ERR_STATE, "plane not available in this state");
        const int rc = be.download(out, src + (size_t)frame * plane, plane * sizeof(float));
        if (rc) return fail(rc, be.last_error());
        if (what == 0 && fold_norm && gmax > 0.f)       // the normalisation lives in the bilateral's patch load
            for (long long i = 0; i < plane; ++i) out[i] = out[i] / gmax;
        return DM_OK;
    }
};
