// Host backend of the engine templates (TEST INFRASTRUCTURE ONLY): every "kernel" launch is a plain loop.
#pragma once
#include <stdlib.h>
#include <string.h>

struct DmHostBackend {
    long long launches = 0;
    const char *last_error() const { return "host backend"; }
    int alloc(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); if (*p) memset(*p, 0xff, bytes); return *p ? 0 : DM_ERR_NOMEM; }
    void free(void *p) { ::free(p); }
    int sync() { return 0; }
    void reset_counters() { launches = 0; }
    void mark_start() {}
    void mark_end() {}
    int upload_small(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
    int upload_frame(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, int rows, int)
    {
        for (int y = 0; y < rows; ++y) memcpy((char *)dst + y * dpitch, (const char *)src + y * spitch, row_bytes);
        return 0;
    }
    int upload_frame_behind(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, int rows)
    {
        return upload_frame(dst, dpitch, src, spitch, row_bytes, rows, 0);
    }
    int wait_uploads() { return 0; }
    int download(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
    int download2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, int rows)
    {
        for (int y = 0; y < rows; ++y) memcpy((char *)dst + y * dpitch, (const char *)src + y * spitch, row_bytes);
        return 0;
    }
#ifdef ECC_NSUMS
    int ecc_search(const EccParams &base, double tx, double ty, int radius, double *, double *host_sums)
    {
        const int count = (2 * radius + 1) * (2 * radius + 1);
        for (int c = 0; c < count; ++c) {
            const EccParams p = ecc_search_candidate(base, tx, ty, radius, c);
            double *out = host_sums + (size_t)c * ECC_NSUMS;
            for (int i = 0; i < ECC_NSUMS; ++i) out[i] = 0;
            float s[ECC_NSUMS];
            for (int y = 0; y < p.h; ++y)
                for (int x = 0; x < p.w; ++x)
                    if (ecc_pixel(p, x, y, s))
                        for (int i = 0; i < ECC_NSUMS; ++i) out[i] += s[i];
        }
        ++launches;
        return 0;
    }
    int ecc_set_state(EccState *dst, const EccState &st) { *dst = st; return 0; }
    int ecc_run_level(const EccLevelArgs &g, EccState *st, double *, unsigned int *)
    {
        if (st->fail) return 0;
        ecc_level_begin(g, *st);
        while (!st->done) {
            const EccParams p = ecc_params_of(g, *st);
            double S[ECC_NSUMS];
            for (int i = 0; i < ECC_NSUMS; ++i) S[i] = 0;
            float s[ECC_NSUMS];
            for (int y = 0; y < p.h; ++y)
                for (int x = 0; x < p.w; ++x)
                    if (ecc_pixel(p, x, y, s))
                        for (int i = 0; i < ECC_NSUMS; ++i) S[i] += s[i];
            ++launches;
            ecc_update(g, *st, S);
        }
        if (!st->fail) ecc_level_end(g, *st);
        return 0;
    }
#endif
    int zero(void *p, size_t bytes) { memset(p, 0, bytes); return 0; }
    template <class K>
    int launch_items(const K &k, const int *count)
    {
        for (int i = 0; i < *count; ++i) k.item(i, 0, 1);
        ++launches;
        return 0;
    }
    template <class K>
    int launch_tiles(const K &k, int tiles_x, int tiles_y, int z, size_t smem_bytes)
    {
        float *sm = (float *)malloc(smem_bytes);
        for (int c = 0; c < z; ++c)
            for (int ty = 0; ty < tiles_y; ++ty)
                for (int tx = 0; tx < tiles_x; ++tx) {
                    memset(sm, 0xff, smem_bytes);       // a read of a never-written slot shows up as NaN
                    k.tile(tx, ty, c, sm, 0, 1);
                }
        ::free(sm);
        ++launches;
        return 0;
    }
    template <class K>
    int launch(const K &k, int w, int h, int z)
    {
        for (int c = 0; c < z; ++c)
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x) k(x, y, c);
        ++launches;
        return 0;
    }
};

