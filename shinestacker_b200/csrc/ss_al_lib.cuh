// ss_al_lib.cuh -- the ssal_* C ABI (include/shinestack_b200_align.h) on the CUDA backend.
// Included at the end of ss_lib.cu after ss_dm_lib.cuh.
#pragma once
#include "../../include/shinestack_b200_align.h"
#include "ss_al_engine.h"

struct ssal_handle {
    AlEngine<DmCudaBackend> eng;
    int device = 0;
};

#define SSAL_ENTER(h)                                                       \
    if (!(h)) return ss_fail(SS_ERR_STATE, "null handle");                  \
    SS_CUDA(cudaSetDevice((h)->device))

SS_EXPORT int ssal_create(int32_t width, int32_t height, int32_t bytes_per_sample, int32_t device, ssal_handle **out)
{
    if (!out) return ss_fail(SS_ERR_INVALID_OPTION, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
        return ss_fail(SS_ERR_NO_DEVICE, "no CUDA device %d (this library has no CPU fallback)", device);
    cudaDeviceProp prop;
    SS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return ss_fail(SS_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                       prop.major, prop.minor);
    SS_CUDA(cudaSetDevice(device));
    ssal_handle *h = new (std::nothrow) ssal_handle();
    if (!h) return ss_fail(SS_ERR_NOMEM, "out of host memory");
    h->device = device;
    int rc = h->eng.be.init();
    if (rc) { ss_fail(rc, "%s", h->eng.be.last_error()); h->eng.be.shutdown(); delete h; return rc; }
    rc = h->eng.create(width, height, bytes_per_sample);
    if (rc) {
        ss_fail(rc, "%s", h->eng.err);
        h->eng.destroy(); h->eng.be.shutdown(); delete h;
        return rc;
    }
    *out = h;
    return SS_OK;
}

SS_EXPORT void ssal_destroy(ssal_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    h->eng.destroy();
    h->eng.be.shutdown();
    delete h;
}

static int ssal_warp_any(ssal_handle *h, const void *src, size_t spitch, int on_device, const double *m, int transform,
                         int border_mode, const double *bval, const int32_t *kernel, int taps, void *dst, size_t dpitch)
{
    SSAL_ENTER(h);
    h->eng.be.reset_timing();
    const int rc = h->eng.warp(src, spitch, on_device, m, transform, border_mode, bval, kernel, taps, dst, dpitch);
    return rc ? ss_fail(rc, "%s", h->eng.err) : SS_OK;
}
SS_EXPORT int ssal_warp(ssal_handle *h, const void *src, size_t spitch, const double *m, int transform, int border_mode,
                        const double *bval, const int32_t *kernel, int taps, void *dst, size_t dpitch)
{
    return ssal_warp_any(h, src, spitch, 0, m, transform, border_mode, bval, kernel, taps, dst, dpitch);
}
SS_EXPORT int ssal_warp_device(ssal_handle *h, const void *src, size_t spitch, const double *m, int transform,
                               int border_mode, const double *bval, const int32_t *kernel, int taps, void *dst,
                               size_t dpitch)
{
    return ssal_warp_any(h, src, spitch, 1, m, transform, border_mode, bval, kernel, taps, dst, dpitch);
}
SS_EXPORT int ssal_ecc_set_reference(ssal_handle *h, const void *ref_bgr, size_t pitch)
{
    SSAL_ENTER(h);
    const int rc = h->eng.ecc_set_reference(ref_bgr, pitch, 0);
    return rc ? ss_fail(rc, "%s", h->eng.err) : SS_OK;
}
SS_EXPORT int ssal_ecc_estimate(ssal_handle *h, const void *mov_bgr, size_t pitch, int finest_level, int max_iters,
                                double eps_px, const double *init_m, double *m_out, double *info)
{
    SSAL_ENTER(h);
    h->eng.be.reset_timing();
    const int rc = h->eng.ecc_estimate(mov_bgr, pitch, 0, finest_level, max_iters, eps_px, init_m, m_out, info);
    return rc ? ss_fail(rc, "%s", h->eng.err) : SS_OK;
}
SS_EXPORT int ssal_prefetch(ssal_handle *h, const void *next_bgr, size_t pitch)
{
    SSAL_ENTER(h);
    const int rc = h->eng.prefetch(next_bgr, pitch);
    return rc ? ss_fail(rc, "%s", h->eng.err) : SS_OK;
}
SS_EXPORT int ssal_use_prefetched(ssal_handle *h)
{
    SSAL_ENTER(h);
    const int rc = h->eng.use_prefetched();
    return rc ? ss_fail(rc, "%s", h->eng.err) : SS_OK;
}
SS_EXPORT int ssal_sync(ssal_handle *h)
{
    SSAL_ENTER(h);
    const int rc = h->eng.be.sync();
    return rc ? ss_fail(rc, "%s", h->eng.be.last_error()) : SS_OK;
}
SS_EXPORT int ssal_result_device(ssal_handle *h, const void **d_bgr, size_t *pitch)
{
    if (!h || !d_bgr || !pitch) return ss_fail(SS_ERR_STATE, "null argument");
    if (!h->eng.result) return ss_fail(SS_ERR_STATE, "no frame has been warped yet");
    *d_bgr = h->eng.result;
    *pitch = h->eng.pitch;
    return SS_OK;
}
SS_EXPORT int ssal_kernel_time_ms(ssal_handle *h, float *ms)
{
    SSAL_ENTER(h);
    const int rc = h->eng.be.elapsed_ms(ms);
    return rc ? ss_fail(rc, "%s", h->eng.be.last_error()) : SS_OK;
}
SS_EXPORT int ssal_launch_count(ssal_handle *h, int64_t *launches)
{
    if (!h || !launches) return ss_fail(SS_ERR_STATE, "null argument");
    *launches = h->eng.be.launches;
    return SS_OK;
}
