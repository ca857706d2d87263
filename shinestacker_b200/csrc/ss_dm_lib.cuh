// ss_dm_lib.cuh -- CUDA backend of DmEngine + the ssdm_* C ABI (include/shinestack_b200_depthmap.h).
// Included at the end of ss_lib.cu (shares its error plumbing); sm_100a only, no CPU fallback.
#pragma once
#include "../../include/shinestack_b200_depthmap.h"
#include "ss_dm_engine.h"
#include "ss_ecc_kernels.cuh"

struct DmCudaBackend {
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ev_copied = nullptr, ev_start = nullptr, ev_end = nullptr;
    bool started = false, ended = false;
    long long launches = 0;
    char msg[256] = "";

    int cuda_fail(cudaError_t e, const char *what)
    {
        snprintf(msg, sizeof(msg), "%s failed: %s", what, cudaGetErrorString(e));
        return DM_ERR_CUDA;
    }
#define DM_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)
    const char *last_error() const { return msg; }
    int init()
    {
        DM_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
        DM_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
        DM_CUDA(cudaEventCreateWithFlags(&ev_copied, cudaEventDisableTiming));
        DM_CUDA(cudaEventCreate(&ev_start));
        DM_CUDA(cudaEventCreate(&ev_end));
        return 0;
    }
    void shutdown()
    {
        if (s_copy) cudaStreamDestroy(s_copy);
        if (s_comp) cudaStreamDestroy(s_comp);
        if (ev_copied) cudaEventDestroy(ev_copied);
        if (ev_start) cudaEventDestroy(ev_start);
        if (ev_end) cudaEventDestroy(ev_end);
        if (h_pinned) cudaFreeHost(h_pinned);
        h_pinned = nullptr; h_pinned_n = 0;
    }
    int alloc(void **p, size_t bytes)
    {
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
        if (e != cudaSuccess) { *p = nullptr; snprintf(msg, sizeof(msg), "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return DM_ERR_NOMEM; }
        return 0;
    }
    void free(void *p) { cudaFree(p); }
    int sync()
    {
        DM_CUDA(cudaStreamSynchronize(s_copy));
        DM_CUDA(cudaStreamSynchronize(s_comp));
        return 0;
    }
    int sync_copies() { DM_CUDA(cudaStreamSynchronize(s_copy)); return 0; }
    void reset_counters() { launches = 0; started = ended = false; }
    void reset_timing() { started = ended = false; }
    void mark_start() { if (!started) { cudaEventRecord(ev_start, s_comp); started = true; ended = false; } }
    void mark_end() { cudaEventRecord(ev_end, s_comp); ended = true; }
    int upload_small(void *dst, const void *src, size_t bytes)
    {
        DM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s_comp));
        DM_CUDA(cudaStreamSynchronize(s_comp));
        return 0;
    }
    // frame -> resident store: host frames go through the copy stream (overlaps the kernels of earlier
    // frames), the compute stream waits for the copy; device frames are copied on the compute stream
    int upload_frame(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, int rows, int on_device)
    {
        if (on_device) {
            DM_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, cudaMemcpyDeviceToDevice, s_comp));
            return 0;
        }
        DM_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, cudaMemcpyHostToDevice, s_copy));
        DM_CUDA(cudaEventRecord(ev_copied, s_copy));
        DM_CUDA(cudaStreamWaitEvent(s_comp, ev_copied, 0));
        return 0;
    }
    // host frame -> device on the copy stream WITHOUT making the compute stream wait; wait_uploads() adds the wait
    int upload_frame_behind(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, int rows)
    {
        DM_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, cudaMemcpyHostToDevice, s_copy));
        DM_CUDA(cudaEventRecord(ev_copied, s_copy));
        return 0;
    }
    int wait_uploads() { DM_CUDA(cudaStreamWaitEvent(s_comp, ev_copied, 0)); return 0; }
    int download(void *dst, const void *src, size_t bytes)
    {
        DM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s_comp));
        DM_CUDA(cudaStreamSynchronize(s_comp));
        return 0;
    }
    int download2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, int rows)
    {
        DM_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, cudaMemcpyDeviceToHost, s_comp));
        DM_CUDA(cudaStreamSynchronize(s_comp));
        return 0;
    }
    template <class K>
    int launch(const K &k, int w, int h, int z)
    {
        const dim3 grid((w + 31) / 32, (h + 7) / 8, z);
        dm_px_kernel<K><<<grid, 256, 0, s_comp>>>(k, w, h);
        DM_CUDA(cudaGetLastError());
        ++launches;
        return 0;
    }
    template <class K>
    int launch_tiles(const K &k, int tiles_x, int tiles_y, int z, size_t smem_bytes)
    {
        static size_t opted = 0;            // per instantiation: largest dynamic shared memory opted in so far
        if (smem_bytes > 48 * 1024 && smem_bytes > opted) {
            DM_CUDA(cudaFuncSetAttribute(dm_tile_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
            opted = smem_bytes;
        }
        dm_tile_kernel<K><<<dim3(tiles_x, tiles_y, z), 256, smem_bytes, s_comp>>>(k);
        DM_CUDA(cudaGetLastError());
        ++launches;
        return 0;
    }
    int zero(void *p, size_t bytes) { DM_CUDA(cudaMemsetAsync(p, 0, bytes, s_comp)); return 0; }
    // one warp per item of a device-side list whose length sits in device memory (persistent grid)
    template <class K>
    int launch_items(const K &k, const int *d_count)
    {
        static int n_sm = 0;
        if (!n_sm) {
            int dev = 0;
            DM_CUDA(cudaGetDevice(&dev));
            DM_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        }
        dm_item_kernel<K><<<8 * n_sm, 256, 0, s_comp>>>(k, d_count);
        DM_CUDA(cudaGetLastError());
        ++launches;
        return 0;
    }
    // ECC passes: `count` parameter sets (one ECC iteration, or all candidates of the coarse search) are
    // evaluated back to back -- zero the count x 28 device sums, one persistent-grid launch per set, ONE
    // download through pinned memory and ONE synchronisation for all of them
    double *h_pinned = nullptr;
    size_t h_pinned_n = 0;
    // Coarse search: the (2 radius + 1)^2 shifted copies of `base` are evaluated by ONE launch (grid.y = candidate,
    // parameters by value: no host-to-device copy), one download through pinned memory, one synchronisation.
    int ecc_search(const EccParams &base, double tx, double ty, int radius, double *d_sums, double *host_sums)
    {
        static int n_sm = 0;
        if (!n_sm) {
            int dev = 0;
            DM_CUDA(cudaGetDevice(&dev));
            DM_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        }
        const int count = (2 * radius + 1) * (2 * radius + 1);
        const size_t n_d = (size_t)count * ECC_NSUMS;
        if (h_pinned_n < n_d) {
            if (h_pinned) cudaFreeHost(h_pinned);
            h_pinned = nullptr; h_pinned_n = 0;
            DM_CUDA(cudaHostAlloc((void **)&h_pinned, sizeof(double) * n_d, cudaHostAllocDefault));
            h_pinned_n = n_d;
        }
        DM_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * n_d, s_comp));
        const long long n = (long long)base.h * base.w;
        long long blocks = (n + 255) / 256;
        if (blocks > 4LL * n_sm) blocks = 4LL * n_sm;     // 4 resident CTAs of 256 threads per SM
        ecc_accumulate_kernel<<<dim3((unsigned)blocks, (unsigned)count), 256, 0, s_comp>>>(base, tx, ty, radius, d_sums);
        ++launches;
        DM_CUDA(cudaGetLastError());
        DM_CUDA(cudaMemcpyAsync(h_pinned, d_sums, sizeof(double) * n_d, cudaMemcpyDeviceToHost, s_comp));
        DM_CUDA(cudaStreamSynchronize(s_comp));
        memcpy(host_sums, h_pinned, sizeof(double) * n_d);
        return 0;
    }
    int ecc_set_state(EccState *d_state, const EccState &st)
    {
        ecc_set_state_kernel<<<1, 1, 0, s_comp>>>(st, d_state);
        DM_CUDA(cudaGetLastError());
        ++launches;
        return 0;
    }
    // All iterations of one pyramid level on the device (cooperative launch; no host round trip).  d_bar: one
    // unsigned int behind the state, the grid barrier's counter.
    int ecc_run_level(const EccLevelArgs &g, EccState *d_state, double *d_sums, unsigned int *d_bar)
    {
        static int n_sm = 0, per_sm = 0;
        if (!n_sm) {
            int dev = 0;
            DM_CUDA(cudaGetDevice(&dev));
            DM_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
            DM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ecc_level_kernel, 256, 0));
            if (per_sm > 4) per_sm = 4;
            if (per_sm < 1) per_sm = 1;
        }
        const long long n = (long long)g.h * g.w;
        long long blocks = (n + 255) / 256;
        if (blocks > (long long)per_sm * n_sm) blocks = (long long)per_sm * n_sm;
        DM_CUDA(cudaMemsetAsync(d_bar, 0, sizeof(unsigned int), s_comp));
        void *args[] = {(void *)&g, (void *)&d_state, (void *)&d_sums, (void *)&d_bar};
        DM_CUDA(cudaLaunchCooperativeKernel((const void *)ecc_level_kernel, dim3((unsigned)blocks), dim3(256), args, 0, s_comp));
        ++launches;
        return 0;
    }
    int elapsed_ms(float *ms)
    {
        *ms = 0.f;
        if (!started || !ended) return 0;
        DM_CUDA(cudaEventSynchronize(ev_end));
        DM_CUDA(cudaEventElapsedTime(ms, ev_start, ev_end));
        return 0;
    }
};

struct ssdm_handle {
    DmEngine<DmCudaBackend> eng;
    int device = 0;
    void *peer_open[DM_MAX_RANKS][18] = {};     // IPC mappings of the peers' exchangeable buffers (ssdm_shard_connect)
};

#define SSDM_ENTER(h)                                                       \
    if (!(h)) return ss_fail(SS_ERR_STATE, "null handle");                  \
    SS_CUDA(cudaSetDevice((h)->device))
#define SSDM_RET(h, rc) do { const int rc_ = (rc); return rc_ ? ss_fail(rc_, "%s", (h)->eng.err) : SS_OK; } while (0)

SS_EXPORT int ssdm_create(const ssdm_config *cfg, ssdm_handle **out)
{
    if (!cfg || !out) return ss_fail(SS_ERR_INVALID_OPTION, "null argument");
    *out = nullptr;
    if (cfg->float_bits != 32)
        return ss_fail(SS_ERR_INVALID_OPTION, "float_type: the device depth-map path implements float-32 only");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
        return ss_fail(SS_ERR_NO_DEVICE, "no CUDA device %d (this library has no CPU fallback)", cfg->device);
    cudaDeviceProp prop;
    SS_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return ss_fail(SS_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device,
                       prop.major, prop.minor);
    SS_CUDA(cudaSetDevice(cfg->device));
    ssdm_handle *h = new (std::nothrow) ssdm_handle();
    if (!h) return ss_fail(SS_ERR_NOMEM, "out of host memory");
    h->device = cfg->device;
    int rc = h->eng.be.init();
    if (rc) { ss_fail(rc, "%s", h->eng.be.last_error()); h->eng.be.shutdown(); delete h; return rc; }
    DmOptions o;
    o.width = cfg->width; o.height = cfg->height; o.bytes_per_sample = cfg->bytes_per_sample;
    o.map_type = cfg->map_type; o.energy = cfg->energy; o.kernel_size = cfg->kernel_size;
    o.blur_size = cfg->blur_size; o.smooth_size = cfg->smooth_size; o.temperature = cfg->temperature;
    o.levels = cfg->levels; o.max_frames = cfg->max_frames;
    rc = h->eng.create(o);
    if (rc) {
        ss_fail(rc, "%s", h->eng.err);
        h->eng.destroy(); h->eng.be.shutdown(); delete h;
        return rc;
    }
    *out = h;
    return SS_OK;
}

static void ssdm_close_peers(ssdm_handle *h)
{
    for (int r = 0; r < DM_MAX_RANKS; ++r)
        for (int k = 0; k < 18; ++k)
            if (h->peer_open[r][k]) { cudaIpcCloseMemHandle(h->peer_open[r][k]); h->peer_open[r][k] = nullptr; }
}

SS_EXPORT void ssdm_destroy(ssdm_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    ssdm_close_peers(h);
    h->eng.destroy();
    h->eng.be.shutdown();
    delete h;
}

SS_EXPORT int ssdm_reset(ssdm_handle *h) { SSDM_ENTER(h); SSDM_RET(h, h->eng.reset()); }
SS_EXPORT int ssdm_push_frame(ssdm_handle *h, const void *bgr, size_t pitch) { SSDM_ENTER(h); SSDM_RET(h, h->eng.push_frame(bgr, pitch, 0)); }
SS_EXPORT int ssdm_push_frame_device(ssdm_handle *h, const void *d_bgr, size_t pitch) { SSDM_ENTER(h); SSDM_RET(h, h->eng.push_frame(d_bgr, pitch, 1)); }
SS_EXPORT int ssdm_sync_copies(ssdm_handle *h)
{
    SSDM_ENTER(h);
    const int rc = h->eng.be.sync_copies();
    return rc ? ss_fail(rc, "%s", h->eng.be.last_error()) : SS_OK;
}
SS_EXPORT int ssdm_finish_begin(ssdm_handle *h) { SSDM_ENTER(h); SSDM_RET(h, h->eng.finish_begin()); }
SS_EXPORT int ssdm_blend_frame(ssdm_handle *h, int frame) { SSDM_ENTER(h); SSDM_RET(h, h->eng.blend_frame(frame)); }
SS_EXPORT int ssdm_finish_end(ssdm_handle *h, void *out_bgr, size_t pitch) { SSDM_ENTER(h); SSDM_RET(h, h->eng.finish_end(out_bgr, pitch)); }
SS_EXPORT int ssdm_finish(ssdm_handle *h, void *out_bgr, size_t pitch)
{
    SSDM_ENTER(h);
    int rc = h->eng.finish_begin();
    for (int i = 0; !rc && i < h->eng.n; ++i) rc = h->eng.blend_frame(i);
    if (!rc) rc = h->eng.finish_end(out_bgr, pitch);
    SSDM_RET(h, rc);
}
SS_EXPORT int ssdm_sync(ssdm_handle *h)
{
    SSDM_ENTER(h);
    const int rc = h->eng.be.sync();
    return rc ? ss_fail(rc, "%s", h->eng.be.last_error()) : SS_OK;
}
// ---- frames of one stack sharded over the GPUs of a box (one process per GPU) -------------------------------------
SS_EXPORT int ssdm_shard_local_max(ssdm_handle *h, float *max_out)
{
    SSDM_ENTER(h);
    if (!max_out) return ss_fail(SS_ERR_STATE, "null argument");
    SSDM_RET(h, h->eng.shard_local_max(max_out));
}
SS_EXPORT int ssdm_shard_finish_local(ssdm_handle *h, float global_max) { SSDM_ENTER(h); SSDM_RET(h, h->eng.shard_finish_local(global_max)); }
SS_EXPORT int ssdm_shard_buffer_count(ssdm_handle *h) { return h ? 2 + h->eng.L : 0; }
SS_EXPORT int ssdm_shard_ptrs(ssdm_handle *h, void **ptrs)
{
    SSDM_ENTER(h);
    if (!ptrs) return ss_fail(SS_ERR_STATE, "null argument");
    SSDM_RET(h, h->eng.shard_ptrs(ptrs));
}
// blob: ssdm_shard_buffer_count() x 64 bytes, the CUDA IPC handles of this rank's exchangeable buffers
SS_EXPORT int ssdm_shard_export(ssdm_handle *h, void *blob)
{
    SSDM_ENTER(h);
    if (!blob) return ss_fail(SS_ERR_STATE, "null argument");
    void *own[18];
    const int rc = h->eng.shard_ptrs(own);
    if (rc) return ss_fail(rc, "%s", h->eng.err);
    for (int k = 0; k < 2 + h->eng.L; ++k) {
        cudaIpcMemHandle_t mh;
        SS_CUDA(cudaIpcGetMemHandle(&mh, own[k]));
        memcpy((char *)blob + 64 * k, &mh, 64);
    }
    return SS_OK;
}
// blobs: world x (ssdm_shard_export blob), in rank order (the own entry is ignored)
SS_EXPORT int ssdm_shard_connect(ssdm_handle *h, const void *blobs, int world, int rank, int root)
{
    SSDM_ENTER(h);
    if (!blobs || world < 1 || world > DM_MAX_RANKS || rank < 0 || rank >= world) return ss_fail(SS_ERR_INVALID_OPTION, "world / rank out of range");
    const int nb = 2 + h->eng.L;
    ssdm_close_peers(h);
    std::vector<void *> all((size_t)world * nb, nullptr);
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        for (int k = 0; k < nb; ++k) {
            cudaIpcMemHandle_t mh;
            memcpy(&mh, (const char *)blobs + ((size_t)r * nb + k) * 64, 64);
            SS_CUDA(cudaIpcOpenMemHandle(&h->peer_open[r][k], mh, cudaIpcMemLazyEnablePeerAccess));
            all[(size_t)r * nb + k] = h->peer_open[r][k];
        }
    }
    SSDM_RET(h, h->eng.shard_connect(all.data(), world, rank, root));
}
// the same with plain device pointers (handles of one process, e.g. tests): ptrs[r * count + k]
SS_EXPORT int ssdm_shard_connect_ptrs(ssdm_handle *h, void *const *ptrs, int world, int rank, int root)
{
    SSDM_ENTER(h);
    if (!ptrs) return ss_fail(SS_ERR_STATE, "null argument");
    SSDM_RET(h, h->eng.shard_connect(ptrs, world, rank, root));
}
SS_EXPORT int ssdm_shard_softmax_rel(ssdm_handle *h) { SSDM_ENTER(h); SSDM_RET(h, h->eng.shard_softmax_rel()); }
SS_EXPORT int ssdm_shard_finish_weights(ssdm_handle *h) { SSDM_ENTER(h); SSDM_RET(h, h->eng.shard_finish_weights()); }
SS_EXPORT int ssdm_shard_reduce_band(ssdm_handle *h) { SSDM_ENTER(h); SSDM_RET(h, h->eng.shard_reduce_band()); }
SS_EXPORT int ssdm_shard_finish_end(ssdm_handle *h, void *out_bgr, size_t pitch) { SSDM_ENTER(h); SSDM_RET(h, h->eng.shard_finish_end(out_bgr, pitch)); }
SS_EXPORT int ssdm_rerun_resident(ssdm_handle *h) { SSDM_ENTER(h); SSDM_RET(h, h->eng.rerun_resident()); }
SS_EXPORT int ssdm_get_plane(ssdm_handle *h, int what, int frame, float *out) { SSDM_ENTER(h); SSDM_RET(h, h->eng.get_plane(what, frame, out)); }
SS_EXPORT int ssdm_kernel_time_ms(ssdm_handle *h, float *ms)
{
    SSDM_ENTER(h);
    const int rc = h->eng.be.elapsed_ms(ms);
    return rc ? ss_fail(rc, "%s", h->eng.be.last_error()) : SS_OK;
}
SS_EXPORT int ssdm_launch_count(ssdm_handle *h, int64_t *launches)
{
    if (!h || !launches) return ss_fail(SS_ERR_STATE, "null argument");
    *launches = h->eng.be.launches;
    return SS_OK;
}
