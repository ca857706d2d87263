// ss_sub_lib.cuh -- the sssub_* C ABI (include/shinestack_b200_subactions.h): BalanceFrames' table look-up and
// Vignetting's radial divide as elementwise passes over a frame in HBM.  Included at the end of ss_lib.cu.
#pragma once
#include "../../include/shinestack_b200_subactions.h"

struct sssub_handle {
    int device = 0, w = 0, h = 0, bps = 1;
    size_t pitch = 0;
    char *frame = nullptr;
    void *lut = nullptr;          // 3 x num_values samples
    cudaStream_t s = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool timed = false;
    long long launches = 0;
};

// One thread per 4 pixels (12 samples): the frame is streamed once, the tables sit in L1/L2.
template <typename T>
__global__ void __launch_bounds__(256) sssub_lut_kernel(T *frame, size_t pitch, int h, int w, const T *lut, int n_lut,
                                                        int num_values)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x4 >= w) return;
    T *px = (T *)((char *)frame + (size_t)y * pitch) + 3 * x4;
    const int n = (w - x4 < 4 ? w - x4 : 4) * 3;
    for (int i = 0; i < n; ++i) {
        const int c = n_lut == 3 ? i % 3 : 0;
        px[i] = lut[(size_t)c * num_values + px[i]];
    }
}

// reference vignetting.py:16-20 in float64: i0 / (1 + exp(min(10, exp(clip(k * (r - r0), -10, 10)))))
__device__ __forceinline__ double sssub_sigmoid(double r, double i0, double k, double r0)
{
    double t = k * (r - r0);
    t = t < -10.0 ? -10.0 : (t > 10.0 ? 10.0 : t);
    double u = exp(t);
    u = u < 10.0 ? u : 10.0;
    return i0 / (1.0 + exp(u));
}

template <typename T>
__global__ void __launch_bounds__(256) sssub_vignette_kernel(T *frame, size_t pitch, int h, int w, double i0, double k,
                                                             double r0, double v0, double max_corr, double threshold,
                                                             double maxv)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    T *px = (T *)((char *)frame + (size_t)y * pitch) + 3 * x;
    const double dx = (double)x - (double)w / 2.0, dy = (double)y - (double)h / 2.0;
    const double r = sqrt(dx * dx + dy * dy);                    // vignetting.py:84-86
    double v = sssub_sigmoid(r, i0, k, r0) / v0;                 // :87
    v = v < 1e-6 ? 1e-6 : (v > 1.0 ? 1.0 : v);
    if (max_corr < 1.0) v = (1.0 - max_corr) + v * max_corr;     // :88-89
    const T b = px[0], g = px[1], rr = px[2];
    T mn = b < g ? b : g;
    mn = mn < rr ? mn : rr;
    if ((double)mn < threshold) v = 1.0;                         // :90-93
    const T in[3] = {b, g, rr};
#pragma unroll
    for (int c = 0; c < 3; ++c) {                                // :96-97
        double q = (double)in[c] / v;
        q = q < 0.0 ? 0.0 : (q > maxv ? maxv : q);
        px[c] = (T)(unsigned int)q;                              // astype: truncation
    }
}

#define SSSUB_ENTER(h)                                                      \
    if (!(h)) return ss_fail(SS_ERR_STATE, "null handle");                  \
    SS_CUDA(cudaSetDevice((h)->device))

SS_EXPORT void sssub_destroy(sssub_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->s) cudaStreamSynchronize(h->s);
    cudaFree(h->frame); cudaFree(h->lut);
    if (h->e0) cudaEventDestroy(h->e0);
    if (h->e1) cudaEventDestroy(h->e1);
    if (h->s) cudaStreamDestroy(h->s);
    delete h;
}

SS_EXPORT int sssub_create(int32_t width, int32_t height, int32_t bytes_per_sample, int32_t device, sssub_handle **out)
{
    if (!out) return ss_fail(SS_ERR_INVALID_OPTION, "null argument");
    *out = nullptr;
    if (bytes_per_sample != 1 && bytes_per_sample != 2)
        return ss_fail(SS_ERR_BIT_DEPTH, "bytes_per_sample must be 1 or 2 (got %d)", bytes_per_sample);
    if (width < 1 || height < 1) return ss_fail(SS_ERR_SHAPE, "image %dx%d too small", width, height);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return ss_fail(SS_ERR_NO_DEVICE, "no CUDA device %d (this library has no CPU fallback)", device);
    }
    int major = 0;
    SS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10)
        return ss_fail(SS_ERR_NO_DEVICE, "device %d has compute capability %d.x; this library is built for sm_100a only",
                       device, major);
    SS_CUDA(cudaSetDevice(device));
    sssub_handle *h = new (std::nothrow) sssub_handle();
    if (!h) return ss_fail(SS_ERR_NOMEM, "out of host memory");
    h->device = device; h->w = width; h->h = height; h->bps = bytes_per_sample;
    h->pitch = ((size_t)width * 3 * bytes_per_sample + 15) / 16 * 16;
    const size_t nv = bytes_per_sample == 1 ? 256 : 65536;
    if (cudaMalloc((void **)&h->frame, h->pitch * height) != cudaSuccess ||
        cudaMalloc(&h->lut, 3 * nv * bytes_per_sample) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->s, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->e0) != cudaSuccess || cudaEventCreate(&h->e1) != cudaSuccess) {
        const int rc = ss_fail(SS_ERR_CUDA, "allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        sssub_destroy(h);
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    *out = h;
    return SS_OK;
}

SS_EXPORT int sssub_upload(sssub_handle *h, const void *bgr, size_t row_pitch_bytes)
{
    SSSUB_ENTER(h);
    const size_t row = (size_t)h->w * 3 * h->bps;
    if (!bgr || row_pitch_bytes < row) return ss_fail(SS_ERR_SHAPE, "bad source / pitch");
    SS_CUDA(cudaMemcpy2DAsync(h->frame, h->pitch, bgr, row_pitch_bytes, row, h->h, cudaMemcpyHostToDevice, h->s));
    h->timed = false;
    return SS_OK;
}

static int sssub_time_begin(sssub_handle *h)
{
    if (!h->timed) { SS_CUDA(cudaEventRecord(h->e0, h->s)); h->timed = true; }
    return 0;
}

SS_EXPORT int sssub_apply_lut(sssub_handle *h, const void *lut, int n_lut)
{
    SSSUB_ENTER(h);
    if (!lut || (n_lut != 1 && n_lut != 3)) return ss_fail(SS_ERR_INVALID_OPTION, "n_lut must be 1 or 3");
    const int nv = h->bps == 1 ? 256 : 65536;
    SS_CUDA(cudaMemcpyAsync(h->lut, lut, (size_t)n_lut * nv * h->bps, cudaMemcpyHostToDevice, h->s));
    SS_CUDA(cudaStreamSynchronize(h->s));        // the table is pageable host memory of the caller
    int rc = sssub_time_begin(h);
    if (rc) return rc;
    dim3 g(((h->w + 3) / 4 + 255) / 256, h->h);
    if (h->bps == 1) sssub_lut_kernel<uint8_t><<<g, 256, 0, h->s>>>((uint8_t *)h->frame, h->pitch, h->h, h->w, (const uint8_t *)h->lut, n_lut, nv);
    else sssub_lut_kernel<uint16_t><<<g, 256, 0, h->s>>>((uint16_t *)h->frame, h->pitch, h->h, h->w, (const uint16_t *)h->lut, n_lut, nv);
    h->launches++;
    SS_CUDA(cudaGetLastError());
    SS_CUDA(cudaEventRecord(h->e1, h->s));
    return SS_OK;
}

SS_EXPORT int sssub_apply_vignette(sssub_handle *h, const double *params, double v0, double max_correction, double threshold)
{
    SSSUB_ENTER(h);
    if (!params) return ss_fail(SS_ERR_INVALID_OPTION, "null parameters");
    int rc = sssub_time_begin(h);
    if (rc) return rc;
    dim3 g((h->w + 255) / 256, h->h);
    const double maxv = h->bps == 1 ? 255.0 : 65535.0;
    if (h->bps == 1)
        sssub_vignette_kernel<uint8_t><<<g, 256, 0, h->s>>>((uint8_t *)h->frame, h->pitch, h->h, h->w, params[0], params[1],
                                                            params[2], v0, max_correction, threshold, maxv);
    else
        sssub_vignette_kernel<uint16_t><<<g, 256, 0, h->s>>>((uint16_t *)h->frame, h->pitch, h->h, h->w, params[0], params[1],
                                                             params[2], v0, max_correction, threshold, maxv);
    h->launches++;
    SS_CUDA(cudaGetLastError());
    SS_CUDA(cudaEventRecord(h->e1, h->s));
    return SS_OK;
}

SS_EXPORT int sssub_sync(sssub_handle *h)
{
    SSSUB_ENTER(h);
    SS_CUDA(cudaStreamSynchronize(h->s));
    return SS_OK;
}

SS_EXPORT int sssub_download(sssub_handle *h, void *bgr, size_t row_pitch_bytes)
{
    SSSUB_ENTER(h);
    const size_t row = (size_t)h->w * 3 * h->bps;
    if (!bgr || row_pitch_bytes < row) return ss_fail(SS_ERR_SHAPE, "bad destination / pitch");
    SS_CUDA(cudaMemcpy2DAsync(bgr, row_pitch_bytes, h->frame, h->pitch, row, h->h, cudaMemcpyDeviceToHost, h->s));
    SS_CUDA(cudaStreamSynchronize(h->s));
    return SS_OK;
}

SS_EXPORT int sssub_device_frame(sssub_handle *h, const void **d_bgr, size_t *pitch_bytes)
{
    SSSUB_ENTER(h);
    if (d_bgr) *d_bgr = h->frame;
    if (pitch_bytes) *pitch_bytes = h->pitch;
    return SS_OK;
}

SS_EXPORT int sssub_kernel_time_ms(sssub_handle *h, float *ms)
{
    SSSUB_ENTER(h);
    if (!ms) return ss_fail(SS_ERR_STATE, "null argument");
    *ms = 0.f;
    if (!h->timed) return SS_OK;
    SS_CUDA(cudaStreamSynchronize(h->s));
    SS_CUDA(cudaEventElapsedTime(ms, h->e0, h->e1));
    return SS_OK;
}

SS_EXPORT int sssub_launch_count(sssub_handle *h, int64_t *launches)
{
    if (!h || !launches) return ss_fail(SS_ERR_STATE, "null argument");
    *launches = h->launches;
    return SS_OK;
}
