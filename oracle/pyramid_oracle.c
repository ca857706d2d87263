/*
 * oracle/pyramid_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked, imported or called by the
 * product path in shinestacker_b200/; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it).
 *
 * A plain-C restatement of the arithmetic of the reference's Laplacian-pyramid focus stacker,
 * lucalista/shinestacker  src/shinestacker/algorithms/pyramid.py  (file:line cited per function).
 * The reference delegates its pixel arithmetic to two third-party wheels that are NOT vendored under
 * /root/reference and are not version-pinned by it (pyproject.toml:19-34): opencv-python (4.13.0 in
 * this image: cv2.filter2D, cv2.cvtColor, cv2.copyMakeBorder) and numpy (2.3.5: argmax, sum, square,
 * astype).  The exact float32 operation order of those calls was pinned empirically and
 * is restated here; tests/test_oracle_vs_reference.py re-checks this file against the live reference
 * and against cv2 itself whenever they are importable, and tests/golden/ holds vectors produced by
 * the live reference (tests/golden/make_golden.py).
 *
 * Build (see oracle/Makefile):  gcc -O2 -mavx2 -mfma -ffp-contract=off -fopenmp -shared -fPIC
 * -ffp-contract=off is REQUIRED: every fused multiply-add below is written explicitly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SO_API __attribute__((visibility("default")))

/* cv2.BORDER_REFLECT101 index map (used by pyramid.py:25 and :70,:88): i<0 -> -i, i>=n -> 2n-2-i,
 * iterated for windows larger than the image (OpenCV borderInterpolate). */
static inline int refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

/* PyramidBase.__init__, pyramid.py:19-22: 1-D generating kernel and its float64 outer product; cv2
 * converts the float64 kernel to float32 before filtering float32 images. */
SO_API void so_make_kernel(double gen_kernel, float *k25)
{
    double k[5] = {0.25 - gen_kernel / 2.0, 0.25, gen_kernel, 0.25, 0.25 - gen_kernel / 2.0};
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j)
            k25[i * 5 + j] = (float)(k[i] * k[j]);
}

/* PyramidBase.convolve, pyramid.py:24-25: cv2.filter2D(img, -1, K, borderType=REFLECT101) on one
 * float32 plane.  25 taps accumulated from +0 in row-major tap order; one fused multiply-add per tap
 * for output columns x < w - (w % 8), separate multiply and add for the last w % 8 columns (the scalar
 * tail of OpenCV's 8-wide loop). */
SO_API void so_filter5x5(const float *src, int h, int w, const float *k25, float *dst)
{
    const int w_vec = w - (w % 8);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const float *rows[5];
        for (int dy = 0; dy < 5; ++dy) rows[dy] = src + (size_t)refl101(y + dy - 2, h) * w;
        float *d = dst + (size_t)y * w;
        int x = 0;
        /* border columns on the left go through the generic path */
        for (; x < 2 && x < w; ++x) {
            float acc = 0.f;
            for (int dy = 0; dy < 5; ++dy)
                for (int dx = 0; dx < 5; ++dx) {
                    float v = rows[dy][refl101(x + dx - 2, w)], k = k25[dy * 5 + dx];
                    if (x < w_vec) acc = __builtin_fmaf(v, k, acc);
                    else acc = acc + v * k;
                }
            d[x] = acc;
        }
        int x_in_end = w - 2 < w_vec ? w - 2 : w_vec;
        for (; x < x_in_end; ++x) {           /* interior, FMA columns (vectorisable) */
            float acc = 0.f;
            for (int dy = 0; dy < 5; ++dy)
                for (int dx = 0; dx < 5; ++dx)
                    acc = __builtin_fmaf(rows[dy][x + dx - 2], k25[dy * 5 + dx], acc);
            d[x] = acc;
        }
        for (; x < w; ++x) {
            float acc = 0.f;
            for (int dy = 0; dy < 5; ++dy)
                for (int dx = 0; dx < 5; ++dx) {
                    float v = rows[dy][refl101(x + dx - 2, w)], k = k25[dy * 5 + dx];
                    if (x < w_vec) acc = __builtin_fmaf(v, k, acc);
                    else acc = acc + v * k;
                }
            d[x] = acc;
        }
    }
}

/* PyramidBase.reduce_layer, pyramid.py:27-32, one plane: convolve(layer)[::2, ::2].  Only the kept
 * (even,even) outputs are evaluated; each is the same 25-tap chain as so_filter5x5 at the full width
 * w (the FMA / tail rule depends on the un-decimated column).  dst is ceil(h/2) x ceil(w/2). */
SO_API void so_reduce(const float *src, int h, int w, const float *k25, float *dst)
{
    const int h2 = (h + 1) / 2, w2 = (w + 1) / 2, w_vec = w - (w % 8);
#pragma omp parallel for schedule(static)
    for (int yo = 0; yo < h2; ++yo) {
        const int y = 2 * yo;
        const float *rows[5];
        for (int dy = 0; dy < 5; ++dy) rows[dy] = src + (size_t)refl101(y + dy - 2, h) * w;
        for (int xo = 0; xo < w2; ++xo) {
            const int x = 2 * xo;
            float acc = 0.f;
            if (x >= 2 && x + 2 < w && x < w_vec) {
                for (int dy = 0; dy < 5; ++dy)
                    for (int dx = 0; dx < 5; ++dx)
                        acc = __builtin_fmaf(rows[dy][x + dx - 2], k25[dy * 5 + dx], acc);
            } else {
                for (int dy = 0; dy < 5; ++dy)
                    for (int dx = 0; dx < 5; ++dx) {
                        float v = rows[dy][refl101(x + dx - 2, w)], k = k25[dy * 5 + dx];
                        if (x < w_vec) acc = __builtin_fmaf(v, k, acc);
                        else acc = acc + v * k;
                    }
            }
            dst[(size_t)yo * w2 + xo] = acc;
        }
    }
}

/* PyramidBase.expand_layer, pyramid.py:34-41, one plane, restated LITERALLY: zero-stuff to 2h x 2w
 * (values at even/even), convolve at that size, multiply by 4.  dst is 2h x 2w. */
SO_API int so_expand_literal(const float *src, int h, int w, const float *k25, float *dst)
{
    const int H = 2 * h, W = 2 * w;
    float *z = (float *)calloc((size_t)H * W, sizeof(float));
    if (!z) return -1;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) z[(size_t)(2 * y) * W + 2 * x] = src[(size_t)y * w + x];
    so_filter5x5(z, H, W, k25, dst);
    for (size_t i = 0; i < (size_t)H * W; ++i) dst[i] = 4.f * dst[i];
    free(z);
    return 0;
}

/* Same result as so_expand_literal, skipping the taps that hit stuffed zeros: fma(0,k,acc) == acc and
 * acc + 0*k == acc exactly (acc starts at +0 and can never become -0), REFLECT101 on the even-sized
 * stuffed image preserves index parity, and the surviving taps keep their row-major order
 * (SURVEY.md §9 items 2-3).  Writes only the top-left out_h x out_w crop (pyramid.py:136-137, :61-62);
 * the tail rule still refers to the un-cropped width 2w. */
SO_API void so_expand(const float *src, int h, int w, const float *k25, float *dst, int out_h,
                      int out_w)
{
    const int H = 2 * h, W = 2 * w, w_vec = W - (W % 8);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < out_h; ++y) {
        for (int x = 0; x < out_w; ++x) {
            float acc = 0.f;
            const int tail = x >= w_vec;
            for (int dy = 0; dy < 5; ++dy) {
                int yy = refl101(y + dy - 2, H);
                if (yy & 1) continue;
                const float *row = src + (size_t)(yy >> 1) * w;
                for (int dx = 0; dx < 5; ++dx) {
                    int xx = refl101(x + dx - 2, W);
                    if (xx & 1) continue;
                    float v = row[xx >> 1], k = k25[dy * 5 + dx];
                    if (tail) acc = acc + v * k;
                    else acc = __builtin_fmaf(v, k, acc);
                }
            }
            dst[(size_t)y * out_w + x] = 4.f * acc;
        }
    }
}

/* cv2.cvtColor(float32 BGR -> GRAY) as used at pyramid.py:49 and :99-101, IPP order pinned in
 * fma(R, .299f, fma(B, .114f, G * .587f)).  Planes in, plane out. */
SO_API void so_gray(const float *b, const float *g, const float *r, size_t n, float *gray)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i)
        gray[i] = __builtin_fmaf(r[i], 0.299f, __builtin_fmaf(b[i], 0.114f, g[i] * 0.587f));
}

/* energy of one Laplacian level of one frame, pyramid.py:49-50: convolve(square(gray(lap))). */
SO_API int so_energy(const float *b, const float *g, const float *r, int h, int w, const float *k25,
                     float *energy)
{
    size_t n = (size_t)h * w;
    float *sq = (float *)malloc(n * sizeof(float));
    if (!sq) return -1;
    so_gray(b, g, r, n, sq);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) sq[i] = sq[i] * sq[i];
    so_filter5x5(sq, h, w, k25, energy);
    free(sq);
    return 0;
}

/* Streaming form of best = np.argmax(energies, axis=0) and fused = sum_i where(best==i, lap_i, 0)
 * (pyramid.py:51-54): first maximum wins => strict '>' in frame order; fused = 0 + lap_winner, which
 * turns -0.0 into +0.0.  first != 0 initialises the state from this frame. */
SO_API void so_select_update(const float *energy, const float *b, const float *g, const float *r,
                             size_t n, int frame_idx, int first, float *best_e, int32_t *best_idx,
                             float *fb, float *fg, float *fr)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        if (first || energy[i] > best_e[i]) {
            best_e[i] = energy[i];
            best_idx[i] = frame_idx;
            fb[i] = 0.f + b[i];
            fg[i] = 0.f + g[i];
            fr[i] = 0.f + r[i];
        }
    }
}

/* numpy's float32 pairwise summation (the add.reduce inner loop used by ndarray.sum() at
 * pyramid.py:77 and :85) for a contiguous run of n values. */
static float np_pairwise_sum_f32(const float *a, int n)
{
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum_f32(a, n2) + np_pairwise_sum_f32(a + n2, n - n2);
    }
}

/* Base-level gray image, pyramid.py:99-101: cvtColor(base.astype(f32), BGR2GRAY).astype(dtype), the
 * cast being C truncation; result held as int32 for both u8 and u16 stacks. */
SO_API void so_base_gray(const float *b, const float *g, const float *r, size_t n, int32_t *gray)
{
    for (size_t i = 0; i < n; ++i) {
        float v = __builtin_fmaf(r[i], 0.299f, __builtin_fmaf(b[i], 0.114f, g[i] * 0.587f));
        gray[i] = (int32_t)v;
    }
}

/* PyramidBase.entropy / area_entropy / get_pad, pyramid.py:66-82, given the per-level table
 * log_p[v] = np.log(probabilities[v]) as float32 (computed by numpy on the host -- numpy's SIMD log is
 * not glibc's logf): per pixel, over the k x k REFLECT101 window flattened row-major,
 * -1 * pairwise_sum( f32(v) * log_p[v] ). */
SO_API void so_entropy_map(const int32_t *gray, int h, int w, int ksize, const float *log_p,
                           float *out)
{
    const int pad = (ksize - 1) / 2, side = 2 * pad + 1, n = side * side;  /* len(arange(-pad, pad+1)), pyramid.py:18,120 */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        float buf[1024];
        for (int x = 0; x < w; ++x) {
            int c = 0;
            for (int dy = -pad; dy <= pad; ++dy)
                for (int dx = -pad; dx <= pad; ++dx) {
                    int v = gray[(size_t)refl101(y + dy, h) * w + refl101(x + dx, w)];
                    buf[c++] = (float)v * log_p[v];
                }
            out[(size_t)y * w + x] = -1.f * np_pairwise_sum_f32(buf, n);
        }
    }
}

/* PyramidBase.deviation / area_deviation, pyramid.py:84-93: mean = float32(exact_sum / n in float64);
 * pairwise_sum( (f32(v) - mean)^2 ) / f32(n). */
SO_API void so_deviation_map(const int32_t *gray, int h, int w, int ksize, float *out)
{
    const int pad = (ksize - 1) / 2, side = 2 * pad + 1, n = side * side;  /* len(arange(-pad, pad+1)), pyramid.py:18,120 */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        float buf[1024];
        int32_t vals[1024];
        for (int x = 0; x < w; ++x) {
            int c = 0;
            int64_t s = 0;
            for (int dy = -pad; dy <= pad; ++dy)
                for (int dx = -pad; dx <= pad; ++dx) {
                    int v = gray[(size_t)refl101(y + dy, h) * w + refl101(x + dx, w)];
                    vals[c++] = v;
                    s += v;
                }
            float mean = (float)((double)s / (double)n);
            for (int i = 0; i < n; ++i) {
                float d = (float)vals[i] - mean;
                buf[i] = d * d;
            }
            out[(size_t)y * w + x] = np_pairwise_sum_f32(buf, n) / (float)n;
        }
    }
}

/* Streaming form of get_fused_base's two arg-maxes (pyramid.py:104-105). */
SO_API void so_argmax_update(const float *val, size_t n, int frame_idx, int first, float *best_v,
                             int32_t *best_idx)
{
    for (size_t i = 0; i < n; ++i)
        if (first || val[i] > best_v[i]) {
            best_v[i] = val[i];
            best_idx[i] = frame_idx;
        }
}

/* get_fused_base, pyramid.py:106-111, for one plane given the winners' base pixels:
 * ((0 + base[best_e]) + base[best_d]) / 2, accumulated in frame order (float32 addition commutes). */
SO_API void so_fuse_base(const float *base_e, const float *base_d, size_t n, float *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = ((0.f + base_e[i]) + base_d[i]) / 2.f;
}

/* One collapse step, pyramid.py:59-63: img = expand(img)[:h,:w] + layer. */
SO_API int so_collapse_step(const float *img, int h_in, int w_in, const float *layer, int h, int w,
                            const float *k25, float *out)
{
    so_expand(img, h_in, w_in, k25, out, h, w);
    size_t n = (size_t)h * w;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = out[i] + layer[i];
    return 0;
}

/* pyramid.py:64 and :179: np.clip(np.abs(img), 0, max).astype(dtype) (C truncation), written into an
 * interleaved H x W x 3 image at channel c. */
SO_API void so_finalize_plane(const float *img, size_t n, int bytes_per_sample, int channel,
                              void *out_interleaved)
{
    const float maxv = bytes_per_sample == 1 ? 255.f : 65535.f;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        float v = fabsf(img[i]);
        v = v < 0.f ? 0.f : (v > maxv ? maxv : v);
        if (bytes_per_sample == 1) ((uint8_t *)out_interleaved)[i * 3 + channel] = (uint8_t)v;
        else ((uint16_t *)out_interleaved)[i * 3 + channel] = (uint16_t)v;
    }
}

/* img.astype(float_type) (pyramid.py:126) + split of the interleaved BGR image into planes. */
SO_API void so_split_planes(const void *img, size_t n, int bytes_per_sample, float *b, float *g,
                            float *r)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        if (bytes_per_sample == 1) {
            const uint8_t *p = (const uint8_t *)img + i * 3;
            b[i] = (float)p[0]; g[i] = (float)p[1]; r[i] = (float)p[2];
        } else {
            const uint16_t *p = (const uint16_t *)img + i * 3;
            b[i] = (float)p[0]; g[i] = (float)p[1]; r[i] = (float)p[2];
        }
    }
}

/* lap = g_l - expand(g_{l+1})[:h,:w]  (pyramid.py:133-138), one plane. */
SO_API void so_laplacian(const float *g, int h, int w, const float *g_next, int h2, int w2,
                         const float *k25, float *lap)
{
    so_expand(g_next, h2, w2, k25, lap, h, w);
    size_t n = (size_t)h * w;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) lap[i] = g[i] - lap[i];
}

SO_API int so_set_threads(int n)
{
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    extern int omp_get_max_threads(void);
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ======================= float_type = 'float-64' (reference base_stack_algo.py:15-16) =======================
 * With float64 images cv2.filter2D accumulates the 25 taps sequentially with separate multiply and add in
 * float64 (no FMA, no tail rule; probed against cv2 4.13.0), using the float64 kernel as given.  Gray,
 * energy and the arg-max stay float32 (pyramid.py:49: lap.astype(np.float32)). */
SO_API void so_make_kernel_f64(double gen_kernel, double *k25)
{
    double k[5] = {0.25 - gen_kernel / 2.0, 0.25, gen_kernel, 0.25, 0.25 - gen_kernel / 2.0};
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) k25[i * 5 + j] = k[i] * k[j];
}

SO_API void so_reduce_f64(const double *src, int h, int w, const double *k25, double *dst)
{
    const int h2 = (h + 1) / 2, w2 = (w + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int yo = 0; yo < h2; ++yo)
        for (int xo = 0; xo < w2; ++xo) {
            double acc = 0.0;
            for (int dy = 0; dy < 5; ++dy)
                for (int dx = 0; dx < 5; ++dx)
                    acc = acc + src[(size_t)refl101(2 * yo + dy - 2, h) * w + refl101(2 * xo + dx - 2, w)] * k25[dy * 5 + dx];
            dst[(size_t)yo * w2 + xo] = acc;
        }
}

SO_API void so_expand_f64(const double *src, int h, int w, const double *k25, double *dst, int out_h, int out_w)
{
    const int H = 2 * h, W = 2 * w;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < out_h; ++y)
        for (int x = 0; x < out_w; ++x) {
            double acc = 0.0;
            for (int dy = 0; dy < 5; ++dy) {
                int yy = refl101(y + dy - 2, H);
                if (yy & 1) continue;
                for (int dx = 0; dx < 5; ++dx) {
                    int xx = refl101(x + dx - 2, W);
                    if (xx & 1) continue;
                    acc = acc + src[(size_t)(yy >> 1) * w + (xx >> 1)] * k25[dy * 5 + dx];
                }
            }
            dst[(size_t)y * out_w + x] = 4.0 * acc;
        }
}

SO_API void so_laplacian_f64(const double *g, int h, int w, const double *g_next, int h2, int w2,
                             const double *k25, double *lap)
{
    so_expand_f64(g_next, h2, w2, k25, lap, h, w);
    size_t n = (size_t)h * w;
    for (size_t i = 0; i < n; ++i) lap[i] = g[i] - lap[i];
}

SO_API void so_collapse_step_f64(const double *img, int h_in, int w_in, const double *layer, int h, int w,
                                 const double *k25, double *out)
{
    so_expand_f64(img, h_in, w_in, k25, out, h, w);
    size_t n = (size_t)h * w;
    for (size_t i = 0; i < n; ++i) out[i] = out[i] + layer[i];
}

SO_API void so_split_planes_f64(const void *img, size_t n, int bytes_per_sample, double *b, double *g, double *r)
{
    for (size_t i = 0; i < n; ++i) {
        if (bytes_per_sample == 1) {
            const uint8_t *p = (const uint8_t *)img + i * 3;
            b[i] = p[0]; g[i] = p[1]; r[i] = p[2];
        } else {
            const uint16_t *p = (const uint16_t *)img + i * 3;
            b[i] = p[0]; g[i] = p[1]; r[i] = p[2];
        }
    }
}

SO_API void so_finalize_plane_f64(const double *img, size_t n, int bytes_per_sample, int channel, void *out_interleaved)
{
    const double maxv = bytes_per_sample == 1 ? 255.0 : 65535.0;
    for (size_t i = 0; i < n; ++i) {
        double v = fabs(img[i]);
        v = v < 0.0 ? 0.0 : (v > maxv ? maxv : v);
        if (bytes_per_sample == 1) ((uint8_t *)out_interleaved)[i * 3 + channel] = (uint8_t)v;
        else ((uint16_t *)out_interleaved)[i * 3 + channel] = (uint16_t)v;
    }
}

static double np_pairwise_sum_f64(const double *a, int n)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum_f64(a, n2) + np_pairwise_sum_f64(a + n2, n - n2);
    }
}

/* entropy / deviation with float_type = float64 (pyramid.py:66-93): same expressions, float64 throughout. */
SO_API void so_entropy_map_f64(const int32_t *gray, int h, int w, int ksize, const double *log_p, double *out)
{
    const int pad = (ksize - 1) / 2, side = 2 * pad + 1, n = side * side;
    for (int y = 0; y < h; ++y) {
        double buf[1024];
        for (int x = 0; x < w; ++x) {
            int c = 0;
            for (int dy = -pad; dy <= pad; ++dy)
                for (int dx = -pad; dx <= pad; ++dx) {
                    int v = gray[(size_t)refl101(y + dy, h) * w + refl101(x + dx, w)];
                    buf[c++] = (double)v * log_p[v];
                }
            out[(size_t)y * w + x] = -1.0 * np_pairwise_sum_f64(buf, n);
        }
    }
}

SO_API void so_deviation_map_f64(const int32_t *gray, int h, int w, int ksize, double *out)
{
    const int pad = (ksize - 1) / 2, side = 2 * pad + 1, n = side * side;
    for (int y = 0; y < h; ++y) {
        double buf[1024];
        int32_t vals[1024];
        for (int x = 0; x < w; ++x) {
            int c = 0;
            int64_t s = 0;
            for (int dy = -pad; dy <= pad; ++dy)
                for (int dx = -pad; dx <= pad; ++dx) {
                    int v = gray[(size_t)refl101(y + dy, h) * w + refl101(x + dx, w)];
                    vals[c++] = v;
                    s += v;
                }
            double mean = (double)s / (double)n;
            for (int i = 0; i < n; ++i) {
                double d = (double)vals[i] - mean;
                buf[i] = d * d;
            }
            out[(size_t)y * w + x] = np_pairwise_sum_f64(buf, n) / (double)n;
        }
    }
}
