// ss_kernels.cuh -- device code of the B200 focus-stacking core (sm_100a).
//
// One body, two builds: nvcc compiles it as CUDA kernels (the product); tests/emu/ compiles the very
// same tile code with g++ (-DSS_HOST_EMU) so the index arithmetic (halos, REFLECT101 fix-ups, tail
// columns, register blocking) can be checked on a CPU-only box.  The emulation is test
// infrastructure; the product never uses it.
//
// Arithmetic contract (what makes the result bit-identical to the reference CPU path):
//   * every 5x5 filter output is ONE sequential chain over the 25 taps in row-major order starting
//     from +0: fused multiply-add for output columns x < w - w%8, separate multiply and add for the
//     last w%8 columns (w = width of the array the reference hands to cv2.filter2D);
//   * expand skips the taps that fall on stuffed zeros (exact), keeps row-major order; the "*4" and the
//     following subtraction / addition are folded into one fma(+-4, acc, x): 4*acc is exact, so the
//     single rounding of the fma equals the reference's two-step result;
//   * gray = fma(R,.299f, fma(B,.114f, G*.587f));  energy = filter(gray*gray);
//   * arg-max over frames = strict '>' in frame order; winner coefficients are stored as 0 + lap.
// Build with -fmad=false: every FMA below is explicit and none may be created by the compiler.
//
// Performance model (B200, measured with tools/microbench.cu): scalar FFMA issues at ~125 of 128
// lanes/clk/SM and the packed FFMA2 form is not faster, so the kernel is bound by FP32 issue slots
// (~75 chain steps per input pixel) and by shared-memory bandwidth (128 B/clk/SM, i.e. one float per
// 4 FMAs).  The fast path therefore register-blocks every stage (>= 3-6 FMAs per float loaded from
// shared memory, 64/128-bit shared loads, fully unrolled tap loops with coefficients in the constant
// bank); tiles that touch an image border or a tail column take the generic path.
#pragma once
#include <stdint.h>

#ifdef SS_HOST_EMU
#include <math.h>
#include <string.h>
#define SS_DEV static inline
#define SS_THREADS(tid) for (int tid = 0; tid < SS_NT; ++tid)
#define SS_SYNC()
#define SS_PT_DECL(type, name, n) type name[SS_NT][n]
#define SS_PT(name) name[tid]
SS_DEV float ss_fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
SS_DEV float ss_mul(float a, float b) { return a * b; }
SS_DEV float ss_add(float a, float b) { return a + b; }
struct ss_f2 { float x, y; };
struct ss_f4 { float x, y, z, w; };
struct ss_u2 { uint32_t x, y; };
struct ss_u4 { uint32_t x, y, z, w; };
#else
#define SS_DEV __device__ __forceinline__
#define SS_THREADS(tid) for (int tid = threadIdx.x, ss_once_ = 1; ss_once_; ss_once_ = 0)
#define SS_SYNC() __syncthreads()
#define SS_PT_DECL(type, name, n) type name[n]
#define SS_PT(name) name
SS_DEV float ss_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SS_DEV float ss_mul(float a, float b) { return __fmul_rn(a, b); }
SS_DEV float ss_add(float a, float b) { return __fadd_rn(a, b); }
typedef float2 ss_f2;
typedef float4 ss_f4;
typedef uint2 ss_u2;
typedef uint4 ss_u4;
#endif

// 8-byte asynchronous global -> shared copy (LDGSTS) and its completion wait
#ifdef SS_HOST_EMU
SS_DEV void ss_cp_async8(void *dst, const void *src) { memcpy(dst, src, 8); }
SS_DEV void ss_cp_async16(void *dst, const void *src) { memcpy(dst, src, 16); }
SS_DEV void ss_cp_async_wait() {}
#else
SS_DEV void ss_cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
SS_DEV void ss_cp_async8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
SS_DEV void ss_cp_async_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
#endif

// TMA (cp.async.bulk.tensor) + mbarrier primitives: one elected thread brings the raw uint8 patch of the
// next frame into shared memory with a single instruction; completion is signalled on an mbarrier.
#ifndef SS_HOST_EMU
SS_DEV unsigned ss_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
SS_DEV void ss_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ss_saddr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
SS_DEV void ss_tma_load_2d(void *dst, const void *tmap, int c0, int c1, uint64_t *bar, unsigned bytes)
{
    // order this CTA's earlier generic-proxy accesses to the destination before the async-proxy writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ss_saddr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(ss_saddr(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(ss_saddr(bar))
                 : "memory");
}
// The same box only as far as L2 (no shared memory involved): issued one frame ahead of the real load, which then
// finds its bytes in L2 instead of waiting for HBM.
SS_DEV void ss_tma_prefetch_2d(const void *tmap, int c0, int c1)
{
#ifndef SS_NO_PREFETCH
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
#else
    (void)tmap; (void)c0; (void)c1;
#endif
}
// 1-D bulk copy global -> shared (16-byte aligned, size a multiple of 16), completion on an mbarrier
SS_DEV void ss_bulk_load(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ss_saddr(dst)), "l"(src), "r"(bytes), "r"(ss_saddr(bar))
                 : "memory");
}
SS_DEV void ss_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ss_saddr(bar)), "r"(bytes) : "memory");
}
SS_DEV void ss_mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tSS_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra SS_DONE;\n\tbra SS_WAIT;\n\tSS_DONE:\n\t}" ::"r"(ss_saddr(bar)), "r"(parity)
                 : "memory");
}
#endif

// Tensor memory (TMEM, 512 columns x 128 lanes x 32 bit per SM) as a per-thread scratchpad: lane 32*(warp%4) + lane of
// the warp, 48 columns per thread (warps 4-7 use columns 48..95).  The Laplacians of a thread's own 4 x 4 block are
// parked there between the Laplacian stage and the arg-max update of the same thread -- data that never crosses
// threads, so it need not occupy shared-memory bandwidth (tcgen05.st / tcgen05.ld run on their own datapath).
#define SS_TM_COLS 128
#ifndef SS_HOST_EMU
SS_DEV void ss_tm_st16(uint32_t taddr, const float *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
                 "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
                 "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
                 "r"(__float_as_uint(v[15]))
                 : "memory");
}
SS_DEV void ss_tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
SS_DEV void ss_tm_ld16(uint32_t taddr, float *v)
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
#endif

SS_DEV ss_f2 ss_mk2(float a, float b) { ss_f2 r; r.x = a; r.y = b; return r; }
SS_DEV ss_f4 ss_mk4(float a, float b, float c, float d) { ss_f4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

// Exact integer -> float conversion of byte k / half-word k of a 32-bit word.  Device: one PRMT builds the
// float 2^23 + v (0x4B000000 | v), one FADD removes the 2^23 -- both full-rate pipes, unlike I2F (quarter rate).
#if defined(SS_HOST_EMU) || defined(SS_CVT_I2F)
SS_DEV float ss_u8_to_f32(uint32_t w, int k) { return (float)((w >> (8 * k)) & 0xffu); }
SS_DEV float ss_u16_to_f32(uint32_t w, int k) { return (float)((w >> (16 * k)) & 0xffffu); }
#else
SS_DEV float ss_u8_to_f32(uint32_t w, int k)
{
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + k)), 8388608.0f);
}
SS_DEV float ss_u16_to_f32(uint32_t w, int k)
{
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, k ? 0x7532u : 0x7510u)), 8388608.0f);
}
#endif

// ---- tile geometry of the fused level kernel --------------------------------------------------------
#define SS_NT 256            // threads per CTA; thread (tx,ty) = (tid&15, tid>>4) owns a 4x4 pixel block
#define SS_TX 64             // output tile (pixels of level l)
#define SS_TY 64
// g_l patch P0: rows Y in [y0-6, y0+70), columns X in [x0-8, x0+72) (the reduce needs x0-6..x0+68; the two
// extra columns on the left make every global row segment 8/16-byte aligned)
#define SS_P0W 80
#define SS_P0H 76
#define SS_P0S 84            // row stride (floats): 84 mod 32 = 20 -> 8 consecutive rows hit 8 distinct 16-byte bank groups
#define SS_P0P (SS_P0H * SS_P0S + 24)  // plane stride (floats), = 8 mod 32: with the reduce stage's lane order
                                       // (6 column blocks x 3 channels x row blocks) the 16-byte shared loads of
                                       // consecutive lanes advance by 48 bytes mod 128 -> conflict-free
#define SS_P0X 8             // column of X = x0 in the patch
#define SS_P0Y 6             // row of Y = y0
// g_{l+1} patch P1: I in [y0/2-2, y0/2+34), J in [x0/2-2, x0/2+34); column jj is stored at jj + 1 so that the
// expand stage's 4-column windows (jj = 2tx+1..2tx+4) are 8-byte aligned
#define SS_P1W 36
#define SS_P1H 36
#ifndef SS_P1S
#define SS_P1S 38            // row stride (floats): even (8-byte aligned windows) and = 6 mod 32, so that the lanes of the
                             // halo-ring strips, which walk down the rows, spread over the banks
#endif
#define SS_P1P (SS_P1H * SS_P1S + 4)
#define SS_P1O 1
// gray^2 patch SQ: rows y in [y0-2, y0+66), columns x in [x0-2, x0+66)
#define SS_SQW 68
#define SS_SQH 68
#define SS_SQS 72
// uint8 frames: the raw bytes of the NEXT frame's patch (76 rows x 240 bytes, dense: it is one TMA box) are
// copied asynchronously into the P1 region
// while the energy stage of the current frame runs (P1 is dead by then)
#define SS_RAWS 240
#define SS_P1REGION 4712   // floats: >= 3 * SS_P1P (4332) and >= 76 * 240 / 4 (4560); also the front part of the
                           // 76 x 496-byte uint16 staging area that continues into SQ
#define SS_P1OFF ((3 * SS_P0P + 31) / 32 * 32)   // P1 / raw staging starts 128-byte aligned (TMA destination)
#define SS_SMEM_FLOATS (SS_P1OFF + SS_P1REGION + SS_SQH * SS_SQS)
#define SS_SMEM_BYTES (SS_SMEM_FLOATS * 4 + 16 + 64)   // + the mbarrier + per-warp maxima of gray^2 (energy-stage skip)

struct SsLevelParams {
    int h, w;            // size of g_l
    int h2, w2;          // size of g_{l+1} = ceil(h/2), ceil(w/2)
    // input g_l: level 0 -> interleaved BGR samples (uint8/uint16), else planar float32
    const void *in;      // first frame of the batch (16-byte aligned)
    long long in_frame_stride;   // bytes between frames (multiple of 16)
    long long in_plane_stride;   // planar only: bytes between channel planes (multiple of 16)
    int in_pitch;                // bytes per row (multiple of 16)
    // output g_{l+1}: planar float32, one slot per frame of the batch
    float *gnext;
    long long gnext_frame_stride;  // floats
    long long gnext_plane_stride;  // floats
    int gnext_pitch;               // floats per row
    // running state of level l (planar, tightly packed h x w)
    float *best_e;
    int32_t *best_idx;
    float *fused;        // 3 planes, plane stride h*w
    int n_frames;        // frames in this batch
    int frame_index0;    // global index of the first frame of the batch
    int init;            // 1: the batch starts the stack (no prior state)
    // frame-split for small levels: blockIdx.z = g handles frames [g*n/G, (g+1)*n/G) with a private state
    // slice (offset g*group_stride elements; planes of `fused` stay h*w apart) that always starts fresh;
    // ss_merge_groups_kernel folds the slices into the running state in frame order.
    int groups;
    long long group_stride;
    float *lap_scratch;  // float32 levels: per-CTA 3 x 64 x 64 floats (own-block Laplacians of the current frame,
                         // written and re-read by the same thread; stays in L2) so that P0 is free for the
                         // next frame's patch while the energy stage runs
    int tma_row0;        // uint8 level 0: row of the batch's first frame in the frame-store tensor map (the store
                         // begins 8 bytes into its allocation, so byte column 8 + 3*(x0-8) is 16-byte aligned)
    float skip_factor;   // > 0: the energy stage of a warp is skipped for a frame when skip_factor * (largest gray^2 the
                         // warp's windows can see) < (smallest running best energy of the warp's pixels); see ss_skip_factor
    float k25[25];
};

// Exactness of the skip: with all 25 taps >= 0 every energy is a chain e = fma(v_24, k_24, ... fma(v_0, k_0, +0)) of
// non-negative terms, so e <= (sum k_i) * max(v) * (1 + 2^-24)^25 < (sum k_i) * (1 + 2^-18) * max(v).  If that bound is
// below the running best of every pixel of the warp, no pixel can satisfy the strict '>' and the stage has no effect.
// Returns 0 (skip disabled) when a tap is negative (gen_kernel > 0.5).
static inline float ss_skip_factor(const float *k25)
{
    double sum = 0.0;
    for (int i = 0; i < 25; ++i) {
        if (!(k25[i] >= 0.f)) return 0.f;
        sum += (double)k25[i];
    }
    float f = (float)(sum * (1.0 + 1.0 / 262144.0));
    while ((double)f < sum * (1.0 + 1.0 / 262144.0)) f = __builtin_nextafterf(f, 2.f * f + 1.f);
    return f;
}

#ifdef SS_HOST_EMU
static long long ss_emu_skipped_threads = 0;   // test statistic: thread-frames whose energy stage was skipped
#endif

SS_DEV int ss_reflect_clamp(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    if (i < 0) i = 0;
    if (i >= n) i = n - 1;
    return i;
}

template <typename InT>
SS_DEV void ss_load_px(const SsLevelParams &p, const char *frame, int y, int x, float &b, float &g,
                       float &r)
{
    if constexpr (sizeof(InT) == 4) {  // planar float32 (levels >= 1)
        const char *s = frame + (long long)y * p.in_pitch + 4ll * x;
        b = *(const float *)s;
        g = *(const float *)(s + p.in_plane_stride);
        r = *(const float *)(s + 2 * p.in_plane_stride);
    } else {                           // interleaved BGR uint8 / uint16 (level 0)
        const InT *s = (const InT *)(frame + (long long)y * p.in_pitch) + 3 * x;
        b = (float)s[0]; g = (float)s[1]; r = (float)s[2];
    }
}

// ---- generic (any border / tail column) per-output helpers ------------------------------------------
SS_DEV float ss_conv25_generic(const float *patch, int stride, const float *k25, bool tail)
{
    float acc = 0.f;
    if (tail) {
#pragma unroll
        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
            for (int dx = 0; dx < 5; ++dx)
                acc = ss_add(acc, ss_mul(patch[dy * stride + dx], k25[dy * 5 + dx]));
    } else {
#pragma unroll
        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) acc = ss_fma(patch[dy * stride + dx], k25[dy * 5 + dx], acc);
    }
    return acc;
}

// Laplacian of one pixel (qq,pp in SQ coordinates, image coordinates y,x), all three channels; returns
// gray^2.
SS_DEV float ss_lap_px_generic(const float *P0, const float *P1, const float *k25, int qq, int pp, int y,
                               int x, int i0, int j0, bool tail, float lap[3])
{
    for (int ch = 0; ch < 3; ++ch) {
        const float *g1 = P1 + ch * SS_P1P;
        float acc = 0.f;
        for (int dy = -2; dy <= 2; ++dy) {
            const int zy = y + dy;
            if (zy & 1) continue;
            const int ii = (zy >> 1) - i0;
            for (int dx = -2; dx <= 2; ++dx) {
                const int zx = x + dx;
                if (zx & 1) continue;
                const int jj = (zx >> 1) - j0;
                const float v = g1[ii * SS_P1S + jj + SS_P1O], k = k25[(dy + 2) * 5 + dx + 2];
                acc = tail ? ss_add(acc, ss_mul(v, k)) : ss_fma(v, k, acc);
            }
        }
        // g - 4*acc with one rounding == the reference's (4*acc, exact) followed by the subtraction
        lap[ch] = ss_fma(-4.f, acc, P0[ch * SS_P0P + (qq + SS_P0Y - 2) * SS_P0S + pp + SS_P0X - 2]);
    }
    const float gray = ss_fma(lap[2], 0.299f, ss_fma(lap[0], 0.114f, ss_mul(lap[1], 0.587f)));
    return ss_mul(gray, gray);
}

// ---- fast-path register-blocked stages ---------------------------------------------------------------
// S1 item: one channel, 3 output rows x 6 output columns of g_{l+1}.  Input rows 2*ii0 .. 2*ii0+8 of the
// P0 patch stream through registers; each row feeds the (up to three) output rows whose 5-row window
// contains it, so every chain still sees its taps in row-major order.
SS_DEV void ss_reduce_block(const float *p0c, int ii0, int jj0, const float *k25, float acc[18])
{
#pragma unroll
    for (int i = 0; i < 18; ++i) acc[i] = 0.f;
    // taps of output jj are patch columns 2*jj + 2 .. 2*jj + 6 (SS_P0X = 8 and j0 = x0/2 - 2).  The row segment
    // 2*jj0 + 2 .. 2*jj0 + 16 is fetched as the 16-byte aligned chunks at 2*jj0 .. 2*jj0 + 15 (jj0 is a multiple of
    // 6, so 2*jj0 floats = 48 bytes * cb) plus one scalar: 5 shared loads per row, and the 8 lanes of a quarter
    // warp (6 column blocks, then the next channel: plane stride = 8 mod 32 floats) hit 8 distinct bank groups.
    const float *base = p0c + (2 * ii0) * SS_P0S + 2 * jj0;
    float vv[2][16];
#define SS_S1_LOAD_ROW(dst, src)                                                                   \
    {                                                                                              \
        const ss_f4 a_ = *(const ss_f4 *)(src), b_ = *(const ss_f4 *)((src) + 4),                  \
                    c_ = *(const ss_f4 *)((src) + 8), d_ = *(const ss_f4 *)((src) + 12);           \
        (dst)[0] = a_.z; (dst)[1] = a_.w; (dst)[2] = b_.x; (dst)[3] = b_.y; (dst)[4] = b_.z;       \
        (dst)[5] = b_.w; (dst)[6] = c_.x; (dst)[7] = c_.y; (dst)[8] = c_.z; (dst)[9] = c_.w;       \
        (dst)[10] = d_.x; (dst)[11] = d_.y; (dst)[12] = d_.z; (dst)[13] = d_.w; (dst)[14] = (src)[16]; \
    }
    SS_S1_LOAD_ROW(vv[0], base)
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        // next row's shared loads are issued before this row's FMAs (latency hidden by ~90 FMAs)
        if (r + 1 < 9) SS_S1_LOAD_ROW(vv[(r + 1) & 1], base + (r + 1) * SS_P0S)
        const float *v = vv[r & 1];
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const int dy = r - 2 * o;
            if (dy < 0 || dy > 4) continue;
#pragma unroll
            for (int c = 0; c < 6; ++c)
#pragma unroll
                for (int dx = 0; dx < 5; ++dx)
                    acc[o * 6 + c] = ss_fma(v[2 * c + dx], k25[dy * 5 + dx], acc[o * 6 + c]);
        }
    }
#undef SS_S1_LOAD_ROW
}

// Expand taps for an N x N pixel block whose top-left pixel has even image coordinates (N = 2 or 4), one
// channel.  win[a][b] holds g_{l+1} rows (q0-2)/2 + a and columns (p0-2)/2 + b, a,b < N/2 + 2; output (r,c)
// uses window rows r/2+{0,1,2} with kernel rows {0,2,4} when r is even and rows (r+1)/2+{0,1} with kernel
// rows {1,3} when r is odd (same for the columns): exactly the taps of the zero-stuffed 5x5 filter that do
// not hit a stuffed zero, in row-major order.
template <int N>
SS_DEV void ss_expand_block(const float (*win)[N / 2 + 2], const float *k25, float *acc)
{
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) {
            float a = 0.f;
#pragma unroll
            for (int ty = 0; ty < 3; ++ty) {
                if ((r & 1) && ty == 2) continue;
                const int wr = (r & 1) ? (r + 1) / 2 + ty : r / 2 + ty;
                const int ky = (r & 1) ? 1 + 2 * ty : 2 * ty;
#pragma unroll
                for (int tx = 0; tx < 3; ++tx) {
                    if ((c & 1) && tx == 2) continue;
                    const int wc = (c & 1) ? (c + 1) / 2 + tx : c / 2 + tx;
                    const int kx = (c & 1) ? 1 + 2 * tx : 2 * tx;
                    a = ss_fma(win[wr][wc], k25[ky * 5 + kx], a);
                }
            }
            acc[r * N + c] = a;
        }
}

// Conversion of 8-pixel units (24 / 48 bytes) into the planar float patch.  item i -> (unit u = i / 76, row
// r = i % 76): consecutive threads handle consecutive patch rows, which the row stride of 84 floats spreads
// over distinct bank groups (conflict-free 16-byte shared stores).
SS_DEV void ss_s0_store_unit(float *P0, int r, int u, const float v[3][8])
{
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float *dst = P0 + ch * SS_P0P + r * SS_P0S + u * 8;
        *(ss_f4 *)dst = ss_mk4(v[ch][0], v[ch][1], v[ch][2], v[ch][3]);
        *(ss_f4 *)(dst + 4) = ss_mk4(v[ch][4], v[ch][5], v[ch][6], v[ch][7]);
    }
}

// ---- S0 helpers (interior tiles) --------------------------------------------------------------------------
// uint8: asynchronous copy of the raw patch bytes (rows y0-6.., byte columns (x0-8)*3 .. +240; 8-byte aligned
// because (x0-8)*3 = 8*k) into the staging area.
SS_DEV void ss_s0_async_u8(const SsLevelParams &p, const char *frame, int x0, int y0, int tid, char *raw)
{
    const int c = tid & 31, r0 = tid >> 5;   // chunk column (30 used), first row; rows advance by 8
    if (c >= 30) return;
    const char *src = frame + (long long)(y0 - SS_P0Y + r0) * p.in_pitch + (x0 - SS_P0X) * 3 + 8 * c;
    char *dst = raw + r0 * SS_RAWS + 8 * c;
#pragma unroll
    for (int k = 0; k < 10; ++k)
        if (r0 + 8 * k < SS_P0H)
            ss_cp_async8(dst + 8 * k * SS_RAWS, src + (long long)(8 * k) * p.in_pitch);
}

// uint16: the raw patch (76 rows x 480 bytes, row stride 496) is staged in the P1 + SQ regions, which are
// both dead between the energy stage of one frame and the reduce stage of the next.
#define SS_RAWS16 496
SS_DEV void ss_s0_async_u16(const SsLevelParams &p, const char *frame, int x0, int y0, int tid, char *raw)
{
    const int c = tid & 31, r0 = tid >> 5;   // 16-byte chunk column (30 used), first row
    if (c >= 30) return;
    const char *src = frame + (long long)(y0 - SS_P0Y + r0) * p.in_pitch + (x0 - SS_P0X) * 6 + 16 * c;
    char *dst = raw + r0 * SS_RAWS16 + 16 * c;
#pragma unroll
    for (int k = 0; k < 10; ++k)
        if (r0 + 8 * k < SS_P0H)
            ss_cp_async16(dst + 8 * k * SS_RAWS16, src + (long long)(8 * k) * p.in_pitch);
}

SS_DEV void ss_s0_convert_u16(float *P0, int tid, const char *raw)
{
    for (int i = tid; i < SS_P0H * 10; i += SS_NT) {
        const int u = i / SS_P0H, r = i - u * SS_P0H;
        const char *src = raw + r * SS_RAWS16 + u * 48;
        const ss_u4 a = *(const ss_u4 *)src, b = *(const ss_u4 *)(src + 16), c = *(const ss_u4 *)(src + 32);
        const uint32_t wd[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        float v[3][8];
#pragma unroll
        for (int k = 0; k < 24; ++k) v[k % 3][k / 3] = ss_u16_to_f32(wd[k >> 1], k & 1);
        ss_s0_store_unit(P0, r, u, v);
    }
}

// float32 planes, prefetch form: the 3 x 76 patch rows are 228 bulk copies of 320 bytes, one per thread, completing
// on the CTA's mbarrier (device) / plain copies (host emulation).
SS_DEV void ss_s0_bulk_f32(const SsLevelParams &p, const char *frame, int x0, int y0, int tid, float *P0, void *mbar)
{
    if (tid >= 3 * SS_P0H) return;
    const int ch = tid / SS_P0H, r = tid - ch * SS_P0H;
    const char *src = frame + ch * p.in_plane_stride + (long long)(y0 - SS_P0Y + r) * p.in_pitch + (x0 - SS_P0X) * 4;
    float *dst = P0 + ch * SS_P0P + r * SS_P0S;
#ifdef SS_HOST_EMU
    (void)mbar;
    memcpy(dst, src, SS_P0W * 4);
#else
    if (tid == 0) ss_mbar_expect_tx((uint64_t *)mbar, 3 * SS_P0H * SS_P0W * 4);
    ss_bulk_load(dst, src, SS_P0W * 4, (uint64_t *)mbar);
#endif
}

// float32 planes: the patch rows are copied straight into P0 (16-byte chunks, all in flight at once).
SS_DEV void ss_s0_async_f32(const SsLevelParams &p, const char *frame, int x0, int y0, int tid, float *P0)
{
    const int c = tid & 31, r0 = tid >> 5;   // 16-byte chunk column (20 used), first row
    if (c >= 20) return;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const char *src = frame + ch * p.in_plane_stride + (long long)(y0 - SS_P0Y + r0) * p.in_pitch +
                          (x0 - SS_P0X) * 4 + 16 * c;
        float *dst = P0 + ch * SS_P0P + r0 * SS_P0S + 4 * c;
#pragma unroll
        for (int k = 0; k < 10; ++k)
            if (r0 + 8 * k < SS_P0H)
                ss_cp_async16(dst + 8 * k * SS_P0S, src + (long long)(8 * k) * p.in_pitch);
    }
}

SS_DEV void ss_s0_convert_u8(float *P0, int tid, const char *raw)
{
    for (int i = tid; i < SS_P0H * 10; i += SS_NT) {
        const int u = i / SS_P0H, r = i - u * SS_P0H;
        const char *src = raw + r * SS_RAWS + u * 24;
        const ss_u2 a = *(const ss_u2 *)src, b = *(const ss_u2 *)(src + 8), c = *(const ss_u2 *)(src + 16);
        const uint32_t wd[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
        float v[3][8];
#pragma unroll
        for (int k = 0; k < 24; ++k) v[k % 3][k / 3] = ss_u8_to_f32(wd[k >> 2], k & 3);
        ss_s0_store_unit(P0, r, u, v);
    }
}



// uint8 interior tiles, S0 in "owner" form: every thread converts the 4 x 4 pixel block it owns in the later stages
// (so its stores over the Laplacians parked there need no barrier after its own energy stage) plus two 4-pixel
// chunks of the 6/8-pixel frame around the 64 x 64 interior (rows 0-5 and 70-75 of the patch: 20 chunks each; rows
// 6-69: chunks 0, 1, 18, 19).  A chunk = 12 raw bytes -> 3 x 4 floats.  Lane -> 3 consecutive words (interior: stride
// 3 words over tx, +16 banks for the second row of blocks of the warp): conflict-free.
SS_DEV void ss_s0_chunk_u8(float *P0, const char *raw, int row, int c4)
{
    const uint32_t *src = (const uint32_t *)(raw + row * SS_RAWS + 12 * c4);
    const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
    float *dst = P0 + row * SS_P0S + 4 * c4;
    *(ss_f4 *)dst = ss_mk4(ss_u8_to_f32(w0, 0), ss_u8_to_f32(w0, 3), ss_u8_to_f32(w1, 2), ss_u8_to_f32(w2, 1));
    *(ss_f4 *)(dst + SS_P0P) = ss_mk4(ss_u8_to_f32(w0, 1), ss_u8_to_f32(w1, 0), ss_u8_to_f32(w1, 3), ss_u8_to_f32(w2, 2));
    *(ss_f4 *)(dst + 2 * SS_P0P) = ss_mk4(ss_u8_to_f32(w0, 2), ss_u8_to_f32(w1, 1), ss_u8_to_f32(w2, 0), ss_u8_to_f32(w2, 3));
}

SS_DEV void ss_s0_owner_u8(float *P0, int tid, const char *raw)
{
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int r = 0; r < 4; ++r) ss_s0_chunk_u8(P0, raw, 4 * ty + SS_P0Y + r, tx + 2);
    {
        const int rr = tid / 20, c4 = tid - rr * 20;                 // tid < 240: the top / bottom six rows
        const int k = tid - 240;                                     // else: left / right chunks of rows 6 ..
        if (tid < 240) ss_s0_chunk_u8(P0, raw, rr < 6 ? rr : SS_TY + rr, c4);
        else ss_s0_chunk_u8(P0, raw, SS_P0Y + (k >> 2), (k & 3) < 2 ? (k & 3) : 16 + (k & 3));
    }
    {
        const int k = tid + 16;                                      // items 256 .. 495 are all of the second kind
        if (k < 256) ss_s0_chunk_u8(P0, raw, SS_P0Y + (k >> 2), (k & 3) < 2 ? (k & 3) : 16 + (k & 3));
    }
}

// S0 for tiles that touch an image border: same 8-pixel units as the interior path whenever a unit lies
// inside the image (vector loads from the REFLECT101-mapped row), per-pixel reflected loads otherwise.
template <typename InT>
SS_DEV void ss_s0_border(const SsLevelParams &p, const char *frame, int x0, int y0, int tid, float *P0)
{
    // three independent items per thread, unrolled: all their global loads are in flight together
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
        const int i = tid + kk * SS_NT;
        if (i >= SS_P0H * 10) break;
        const int u = i / SS_P0H, r = i - u * SS_P0H;
        const int y = ss_reflect_clamp(y0 - SS_P0Y + r, p.h);
        const int X = x0 - SS_P0X + 8 * u;
        float v[3][8];
        if (X >= 0 && X + 8 <= p.w) {
            if constexpr (sizeof(InT) == 1) {
                const char *src = frame + (long long)y * p.in_pitch + X * 3;
                const ss_u2 a = *(const ss_u2 *)src, b = *(const ss_u2 *)(src + 8), c = *(const ss_u2 *)(src + 16);
                const uint32_t wd[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
#pragma unroll
                for (int k = 0; k < 24; ++k) v[k % 3][k / 3] = ss_u8_to_f32(wd[k >> 2], k & 3);
            } else if constexpr (sizeof(InT) == 2) {
                const char *src = frame + (long long)y * p.in_pitch + X * 6;
                const ss_u4 a = *(const ss_u4 *)src, b = *(const ss_u4 *)(src + 16), c = *(const ss_u4 *)(src + 32);
                const uint32_t wd[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
                for (int k = 0; k < 24; ++k) v[k % 3][k / 3] = ss_u16_to_f32(wd[k >> 1], k & 1);
            } else {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const char *src = frame + ch * p.in_plane_stride + (long long)y * p.in_pitch + X * 4;
                    const ss_f4 a = *(const ss_f4 *)src, b = *(const ss_f4 *)(src + 16);
                    v[ch][0] = a.x; v[ch][1] = a.y; v[ch][2] = a.z; v[ch][3] = a.w;
                    v[ch][4] = b.x; v[ch][5] = b.y; v[ch][6] = b.z; v[ch][7] = b.w;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                ss_load_px<InT>(p, frame, y, ss_reflect_clamp(X + k, p.w), v[0][k], v[1][k], v[2][k]);
        }
        ss_s0_store_unit(P0, r, u, v);
    }
}

// The fused level step for one tile: for every frame of the batch
//   S0  g_l patch (REFLECT101 at the image border)                        -> smem P0
//   S1  g_{l+1} = reduce(g_l) on the patch (reference pyramid.py:27-32)   -> smem P1 (interior -> global)
//   S2  lap = g_l - 4*expand(g_{l+1}) (pyramid.py:34-46,133-138) parked in place of g_l in P0,
//       gray^2 (pyramid.py:49-50)                                         -> smem SQ
//   S3  energy = filter(gray^2); running strict-'>' arg-max (pyramid.py:50-54): best energy in registers,
//       winner index / coefficients (0 + lap) written to global memory when a pixel's best changes
template <typename InT>
#ifndef SS_HOST_EMU
__device__ __forceinline__
#else
static inline
#endif
void ss_level_tile(const SsLevelParams &p, const void *tmap, int tile_x, int tile_y, int group, int cta_linear,
                   float *smem)
{
    float *P0 = smem;
    float *P1 = P0 + SS_P1OFF;
    float *SQ = P1 + SS_P1REGION;
    const int x0 = tile_x * SS_TX, y0 = tile_y * SS_TY;
    const int w_vec = p.w - (p.w % 8);                // FMA columns of the reduce / energy filters
    const int wz = 2 * p.w2, wz_vec = wz - (wz % 8);  // ... of the expand filter (width 2*w2)
    const int j0 = x0 / 2 - 2, i0 = y0 / 2 - 2;       // g_{l+1} coordinates of P1[0][0]
    // every patch position inside the images?  (else REFLECT101 loads + fix-up passes)
    const bool interior = (x0 >= SS_P0X) && (y0 >= SS_P0Y) && (x0 + SS_TX + 8 <= p.w) &&
                          (y0 + SS_TY + 6 <= p.h) && (x0 / 2 + SS_TX / 2 + 2 <= p.w2) &&
                          (y0 / 2 + SS_TY / 2 + 2 <= p.h2) &&
                          // uint8: the TMA box [8 + 3*(x0-8), +240) must lie inside one row of the tensor map
                          (sizeof(InT) != 1 || 3 * (x0 + SS_TX + 8) + 8 <= p.in_pitch);
    // no IN-IMAGE output of this tile in any stage falls on a tail (mul+add) column?  (An image whose widths
    // are multiples of 8 has no tail columns at all.)  Then the register-blocked stages are used; at image
    // borders they also evaluate out-of-image positions (finite junk computed from clamped data) which the
    // fix-up passes overwrite or which are never stored.
    const bool fast = (w_vec == p.w && wz_vec == wz) || ((x0 + SS_TX + 4 <= w_vec) && (x0 + SS_TX + 4 <= wz_vec));
    const bool full = (x0 + SS_TX <= p.w) && (y0 + SS_TY <= p.h);
    const long long plane_px = (long long)p.h * p.w;
    const int f_begin = p.groups > 1 ? (int)((long long)group * p.n_frames / p.groups) : 0;
    const int f_end = p.groups > 1 ? (int)((long long)(group + 1) * p.n_frames / p.groups) : p.n_frames;
    const bool fresh = p.init || p.groups > 1;
    float *const st_e = p.best_e + group * p.group_stride;
    int32_t *const st_idx = p.best_idx + group * p.group_stride;
    float *const st_fused = p.fused + 3 * group * p.group_stride;
    // rows of the coefficient planes 16-byte aligned?  (vector stores of the first frame's coefficients)
    // g_{l+1} rows 8-byte aligned?  (pair stores from the reduce stage; the base level is tightly packed)
    const bool gpair = (((p.gnext_pitch | p.gnext_plane_stride | p.gnext_frame_stride) & 1) == 0) &&
                       ((((unsigned long long)p.gnext) & 7) == 0);
    const bool state_vec4 = ((p.w & 3) == 0) && ((plane_px & 3) == 0) && ((((unsigned long long)st_fused) & 15) == 0);

    SS_PT_DECL(float, be, 16);
    SS_PT_DECL(float, bemin, 1);   // smallest of the 16 (energy-stage skip)
    // exact skip of the energy stage (see ss_skip_factor): register-blocked tiles that lie fully inside the image
    const bool skip_on = p.skip_factor > 0.f && fast && full;
    float *const wmax = smem + SS_SMEM_FLOATS + 4;   // 8 + 5 floats behind the mbarrier

    // register-blocked tiles park the own-block Laplacians in tensor memory (per-thread data), so P0 is dead after S2;
    // float32 levels, interior fast tiles: the NEXT frame's patch is bulk-copied into P0 while the energy stage runs
    const bool pref32 = sizeof(InT) == 4 && interior && fast;
    (void)cta_linear;
#ifdef SS_HOST_EMU
    SS_PT_DECL(float, tmv, 48);    // stands in for the thread's 48 tensor-memory columns
#else
    uint32_t *tm_slot = (uint32_t *)(smem + SS_SMEM_FLOATS + 4 + 14);
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ss_saddr(tm_slot)), "n"(SS_TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_base = *tm_slot;
    const uint32_t tm_addr = tm_base + ((((threadIdx.x >> 5) & 3) * 32u) << 16) + ((threadIdx.x >> 7) ? 48u : 0u);
#endif
    // uint8, interior tiles: raw patch of frame f arrives by TMA (device) / memcpy (host emulation)
#ifndef SS_HOST_EMU
    uint64_t *mbar = (uint64_t *)(smem + SS_SMEM_FLOATS);
    unsigned mbar_phase = 0;
    if (threadIdx.x == 0) ss_mbar_init(mbar, 1);   // (only the TMA / bulk-copy paths wait on it)
    __syncthreads();
    if (pref32) {
        ss_s0_bulk_f32(p, (const char *)p.in + (long long)f_begin * p.in_frame_stride, x0, y0, threadIdx.x, P0, mbar);
        ss_mbar_wait(mbar, mbar_phase);
        mbar_phase ^= 1;
    }
    if (interior && sizeof(InT) == 1) {
        if (threadIdx.x == 0) {
            ss_tma_load_2d(P1, tmap, 8 + (x0 - SS_P0X) * 3, p.tma_row0 + f_begin * p.h + y0 - SS_P0Y, mbar, SS_P0H * SS_RAWS);
            if (f_begin + 1 < f_end)
                ss_tma_prefetch_2d(tmap, 8 + (x0 - SS_P0X) * 3, p.tma_row0 + (f_begin + 1) * p.h + y0 - SS_P0Y);
        }
        ss_mbar_wait(mbar, mbar_phase);
        mbar_phase ^= 1;
        ss_s0_owner_u8(P0, threadIdx.x, (const char *)P1);
        __syncthreads();
    }
#else
    (void)tmap;
    void *mbar = nullptr;
    if (pref32) {
        SS_THREADS(tid)
        {
            ss_s0_bulk_f32(p, (const char *)p.in + (long long)f_begin * p.in_frame_stride, x0, y0, tid, P0, mbar);
        }
    }
    if (interior && sizeof(InT) == 1) {
        SS_THREADS(tid)
        {
            ss_s0_async_u8(p, (const char *)p.in + (long long)f_begin * p.in_frame_stride, x0, y0, tid, (char *)P1);
            ss_cp_async_wait();
        }
        SS_SYNC();
        SS_THREADS(tid) { ss_s0_owner_u8(P0, tid, (const char *)P1); }
        SS_SYNC();
    }
#endif
    // uint8 interior tiles: the float patch of frame f is produced at the END of frame f-1's energy stage (each thread
    // converts the block it owns + a share of the frame around the interior), so the loop body starts with S1
    const bool s0_owner = interior && sizeof(InT) == 1;
    if (!fresh) {
        SS_THREADS(tid)
        {
            const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int x = x0 + 4 * tx + c, y = y0 + 4 * ty + r;
                    if (x < p.w && y < p.h) SS_PT(be)[r * 4 + c] = st_e[(long long)y * p.w + x];
                }
            if (skip_on) {
                float m = SS_PT(be)[0];
#pragma unroll
                for (int q = 1; q < 16; ++q) m = fminf(m, SS_PT(be)[q]);
                SS_PT(bemin)[0] = m;
            }
        }
    }

    for (int f = f_begin; f < f_end; ++f) {
        const char *frame = (const char *)p.in + (long long)f * p.in_frame_stride;
        const bool first_frame = fresh && f == f_begin;   // every in-image pixel takes this frame
        // ---- S0: g_l patch -------------------------------------------------------------------------
        if (s0_owner) {
            // nothing to do: P0 was filled before the loop / at the end of the previous iteration
        } else if (interior) {
            SS_THREADS(tid)
            {
                if constexpr (sizeof(InT) == 2) {
                    ss_s0_async_u16(p, frame, x0, y0, tid, (char *)P1);
                    ss_cp_async_wait();
                } else if (!pref32) {
                    ss_s0_async_f32(p, frame, x0, y0, tid, P0);
                    ss_cp_async_wait();
                }
            }
            if constexpr (sizeof(InT) == 2) {
                SS_SYNC();
                SS_THREADS(tid) { ss_s0_convert_u16(P0, tid, (const char *)P1); }
            }
        } else {
            SS_THREADS(tid) { ss_s0_border<InT>(p, frame, x0, y0, tid, P0); }
        }
        if (!s0_owner) SS_SYNC();
        // ---- S1: reduce -> P1 ------------------------------------------------------------------------
        if (fast) {
            SS_THREADS(tid)
            {
                if (tid < 216) {
                    const int cb = tid % 6, t2 = tid / 6, ch = t2 % 3, rb = t2 / 3;
                    const int ii0 = 3 * rb, jj0 = 6 * cb;
                    float acc[18];
                    ss_reduce_block(P0 + ch * SS_P0P, ii0, jj0, p.k25, acc);
                    float *dst = P1 + ch * SS_P1P + ii0 * SS_P1S + jj0 + SS_P1O;
#pragma unroll
                    for (int o = 0; o < 3; ++o)
#pragma unroll
                        for (int c = 0; c < 6; ++c) dst[o * SS_P1S + c] = acc[o * 6 + c];
                    // the 32 x 32 interior of P1 is this tile's share of g_{l+1}: straight from the registers to
                    // global memory as aligned pairs (J is even: j0, jj0 and c are).  Row / pair validity is hoisted
                    // into six predicates so that the stores compile to predicated STG.64 without branches.
                    float *gdst = p.gnext + f * p.gnext_frame_stride + ch * p.gnext_plane_stride +
                                  (long long)(i0 + ii0) * p.gnext_pitch + j0 + jj0;
                    bool rok[3], cok[3];
#pragma unroll
                    for (int o = 0; o < 3; ++o)
                        rok[o] = (ii0 + o >= 2) & (ii0 + o < SS_P1H - 2) & (i0 + ii0 + o < p.h2);
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        cok[c] = (jj0 + 2 * c >= 2) & (jj0 + 2 * c < SS_P1W - 2) & (j0 + jj0 + 2 * c + 1 < p.w2);
                    if (gpair) {
#pragma unroll
                        for (int o = 0; o < 3; ++o)
#pragma unroll
                            for (int c = 0; c < 3; ++c)
                                if (rok[o] & cok[c])
                                    *(ss_f2 *)(gdst + o * p.gnext_pitch + 2 * c) = ss_mk2(acc[o * 6 + 2 * c], acc[o * 6 + 2 * c + 1]);
                    } else {
#pragma unroll
                        for (int o = 0; o < 3; ++o)
#pragma unroll
                            for (int c = 0; c < 3; ++c)
                                if (rok[o] & cok[c]) {
                                    gdst[o * p.gnext_pitch + 2 * c] = acc[o * 6 + 2 * c];
                                    gdst[o * p.gnext_pitch + 2 * c + 1] = acc[o * 6 + 2 * c + 1];
                                }
                    }
                    // (fast tiles never contain the last column of an odd-width g_{l+1}: either 2*w2 is a multiple of
                    // 8, or the tile ends at least 4 columns before the tail columns -- so no pair straddles w2)
                }
            }
        } else {
            SS_THREADS(tid)
            {
                for (int i = tid; i < 3 * SS_P1H * SS_P1W; i += SS_NT) {
                    const int ch = i / (SS_P1H * SS_P1W), rem = i - ch * (SS_P1H * SS_P1W);
                    const int ii = rem / SS_P1W, jj = rem - ii * SS_P1W;
                    const int I = i0 + ii, J = j0 + jj;
                    if (I < 0 || I >= p.h2 || J < 0 || J >= p.w2) continue;
                    P1[ch * SS_P1P + ii * SS_P1S + jj + SS_P1O] =
                        ss_conv25_generic(P0 + ch * SS_P0P + 2 * ii * SS_P0S + 2 * jj + 2, SS_P0S, p.k25, 2 * J >= w_vec);
                }
            }
        }
        SS_SYNC();
        if (!interior) {
            // out-of-image g_{l+1} positions as the expand filter sees them: the stuffed image has size
            // 2*h2 x 2*w2 and REFLECT101 on it maps even index -2 -> 2 (J=-1 -> 1) and 2*w2 -> 2*w2-2
            // (J=w2 -> w2-1).
            SS_THREADS(tid)
            {
                for (int i = tid; i < 3 * SS_P1H * SS_P1W; i += SS_NT) {
                    const int ch = i / (SS_P1H * SS_P1W), rem = i - ch * (SS_P1H * SS_P1W);
                    const int ii = rem / SS_P1W, jj = rem - ii * SS_P1W;
                    const int I = i0 + ii, J = j0 + jj;
                    if (I >= 0 && I < p.h2 && J >= 0 && J < p.w2) continue;
                    const int Is = I < 0 ? -I : (I >= p.h2 ? 2 * p.h2 - 1 - I : I);
                    const int Js = J < 0 ? -J : (J >= p.w2 ? 2 * p.w2 - 1 - J : J);
                    if (Is < 0 || Is >= p.h2 || Js < 0 || Js >= p.w2) continue;  // only junk needs it
                    const int is = Is - i0, js = Js - j0;
                    if (is < 0 || is >= SS_P1H || js < 0 || js >= SS_P1W) continue;
                    P1[ch * SS_P1P + ii * SS_P1S + jj + SS_P1O] = P1[ch * SS_P1P + is * SS_P1S + js + SS_P1O];
                }
            }
            SS_SYNC();
        }
        // ---- S2: g_{l+1} interior -> global; Laplacian (parked in P0), gray^2 -> SQ ----------------------
        if (!fast) SS_THREADS(tid)
        {
            // the 32 x 32 interior of P1 is this tile's share of g_{l+1}: thread -> (row tid>>3, 4 columns
            // 4*(tid&7)..) of every channel, written as one 16-byte store when the row pitch allows it
            const int ii = 2 + (tid >> 3), jj = 2 + 4 * (tid & 7);
            const int I = i0 + ii, J = j0 + jj;
            float *gdst = p.gnext + f * p.gnext_frame_stride + (long long)I * p.gnext_pitch + J;
            const float *gsrc = P1 + ii * SS_P1S + jj + SS_P1O;
            if (I < p.h2) {
                if ((p.gnext_pitch & 3) == 0 && J + 4 <= p.w2) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        *(ss_f4 *)(gdst + ch * p.gnext_plane_stride) =
                            ss_mk4(gsrc[ch * SS_P1P], gsrc[ch * SS_P1P + 1], gsrc[ch * SS_P1P + 2], gsrc[ch * SS_P1P + 3]);
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (J + c < p.w2) gdst[ch * p.gnext_plane_stride + c] = gsrc[ch * SS_P1P + c];
                }
            }
        }
        if (fast) {
            SS_THREADS(tid)
            {
                const int tx = tid & 15, ty = tid >> 4;
                // own 4x4 block; gray = fma(R,.299, fma(B,.114, G*.587)) is accumulated in the
                // order G, B, R starting from +0 (fma(G,.587,+0) == G*.587 up to the sign of zero, which the
                // square removes).
                float gray[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) gray[q] = 0.f;
                // channel order G, B, R = planes 1, 0, 2: the plane pointers advance by -1 and then +2 planes
                const float *p1c = P1 + SS_P1P + (2 * ty + 1) * SS_P1S + 2 * tx + 1 + SS_P1O;
                float *p0c = P0 + SS_P0P + (4 * ty + SS_P0Y) * SS_P0S + 4 * tx + SS_P0X;
#pragma unroll
                for (int it = 0; it < 3; ++it) {
                    const float wgt = it == 0 ? 0.587f : (it == 1 ? 0.114f : 0.299f);
                    float win[4][4], acc[16];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const ss_f2 u = *(const ss_f2 *)(p1c + a * SS_P1S), v = *(const ss_f2 *)(p1c + a * SS_P1S + 2);
                        win[a][0] = u.x; win[a][1] = u.y; win[a][2] = v.x; win[a][3] = v.y;
                    }
                    ss_expand_block<4>(win, p.k25, acc);
                    const int chi = it == 0 ? 1 : (it == 1 ? 0 : 2);   // plane of this iteration
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const ss_f4 g = *(const ss_f4 *)(p0c + r * SS_P0S);
                        // the Laplacians replace the expand sums in their registers
                        acc[r * 4 + 0] = ss_fma(-4.f, acc[r * 4 + 0], g.x); acc[r * 4 + 1] = ss_fma(-4.f, acc[r * 4 + 1], g.y);
                        acc[r * 4 + 2] = ss_fma(-4.f, acc[r * 4 + 2], g.z); acc[r * 4 + 3] = ss_fma(-4.f, acc[r * 4 + 3], g.w);
                        if (pref32 && first_frame) {   // all pixels win: straight into the running state (0 + lap)
                            float *fd = st_fused + chi * plane_px + (long long)(y0 + 4 * ty + r) * p.w + x0 + 4 * tx;
                            if (state_vec4) {
                                *(ss_f4 *)fd = ss_mk4(ss_add(0.f, acc[r * 4]), ss_add(0.f, acc[r * 4 + 1]),
                                                      ss_add(0.f, acc[r * 4 + 2]), ss_add(0.f, acc[r * 4 + 3]));
                            } else {
                                fd[0] = ss_add(0.f, acc[r * 4]); fd[1] = ss_add(0.f, acc[r * 4 + 1]);
                                fd[2] = ss_add(0.f, acc[r * 4 + 2]); fd[3] = ss_add(0.f, acc[r * 4 + 3]);
                            }
                        }
                        gray[r * 4 + 0] = ss_fma(acc[r * 4 + 0], wgt, gray[r * 4 + 0]);
                        gray[r * 4 + 1] = ss_fma(acc[r * 4 + 1], wgt, gray[r * 4 + 1]);
                        gray[r * 4 + 2] = ss_fma(acc[r * 4 + 2], wgt, gray[r * 4 + 2]);
                        gray[r * 4 + 3] = ss_fma(acc[r * 4 + 3], wgt, gray[r * 4 + 3]);
                    }
                    // park the block's Laplacians of this plane in the thread's tensor-memory columns
                    if (!(pref32 && first_frame)) {
#ifdef SS_HOST_EMU
                        for (int q = 0; q < 16; ++q) SS_PT(tmv)[chi * 16 + q] = acc[q];
#else
                        ss_tm_st16(tm_addr + chi * 16, acc);
#endif
                    }
                    const int step = it == 0 ? -1 : 2;
                    p1c += step * SS_P1P;
                    p0c += step * SS_P0P;
                }
                float sqmax = 0.f;   // largest gray^2 of the block (squares are >= +0)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float *dst = SQ + (4 * ty + r + 2) * SS_SQS + 4 * tx + 2;
                    const float s0 = ss_mul(gray[r * 4], gray[r * 4]), s1 = ss_mul(gray[r * 4 + 1], gray[r * 4 + 1]);
                    const float s2 = ss_mul(gray[r * 4 + 2], gray[r * 4 + 2]), s3 = ss_mul(gray[r * 4 + 3], gray[r * 4 + 3]);
                    *(ss_f2 *)dst = ss_mk2(s0, s1);
                    *(ss_f2 *)(dst + 2) = ss_mk2(s2, s3);
                    if (skip_on) sqmax = fmaxf(fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)), sqmax);
                }
#ifndef SS_HOST_EMU
                ss_tm_wait_st();
#endif
                float ringmax = 0.f;
                // halo ring of SQ (2 pixels wide): 132 aligned 2x2 quads
                if (tid < 132) {
                    int qq, pp;
                    if (tid < 34) { qq = 0; pp = 2 * tid; }                         // top strip
                    else if (tid < 68) { qq = SS_SQH - 2; pp = 2 * (tid - 34); }     // bottom strip
                    else if (tid < 100) { qq = 2 + 2 * (tid - 68); pp = 0; }         // left strip
                    else { qq = 2 + 2 * (tid - 100); pp = SS_SQW - 2; }              // right strip
                    float gr[4] = {0.f, 0.f, 0.f, 0.f};
                    const float *r1c = P1 + SS_P1P + (qq / 2) * SS_P1S + pp / 2 + SS_P1O;
                    const float *r0c = P0 + SS_P0P + (qq + SS_P0Y - 2) * SS_P0S + pp + SS_P0X - 2;
#pragma unroll
                    for (int it = 0; it < 3; ++it) {
                        const float wgt = it == 0 ? 0.587f : (it == 1 ? 0.114f : 0.299f);
                        float win[3][3], acc[4];
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int b = 0; b < 3; ++b) win[a][b] = r1c[a * SS_P1S + b];
                        ss_expand_block<2>(win, p.k25, acc);
                        // g_l of the quad: two aligned pairs (pp is even)
                        const ss_f2 g0 = *(const ss_f2 *)r0c, g1 = *(const ss_f2 *)(r0c + SS_P0S);
                        gr[0] = ss_fma(ss_fma(-4.f, acc[0], g0.x), wgt, gr[0]);
                        gr[1] = ss_fma(ss_fma(-4.f, acc[1], g0.y), wgt, gr[1]);
                        gr[2] = ss_fma(ss_fma(-4.f, acc[2], g1.x), wgt, gr[2]);
                        gr[3] = ss_fma(ss_fma(-4.f, acc[3], g1.y), wgt, gr[3]);
                        const int step = it == 0 ? -1 : 2;
                        r1c += step * SS_P1P;
                        r0c += step * SS_P0P;
                    }
                    const float s0 = ss_mul(gr[0], gr[0]), s1 = ss_mul(gr[1], gr[1]), s2 = ss_mul(gr[2], gr[2]), s3 = ss_mul(gr[3], gr[3]);
                    *(ss_f2 *)(SQ + qq * SS_SQS + pp) = ss_mk2(s0, s1);
                    *(ss_f2 *)(SQ + (qq + 1) * SS_SQS + pp) = ss_mk2(s2, s3);
                    ringmax = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
                }
                if (skip_on) {
                    // per-warp maxima: wmax[w] = own blocks (rows 8w .. 8w+7 of the tile), wmax[8 + w] = the warp's ring quads
                    const int wp = tid >> 5;
#ifdef SS_HOST_EMU
                    if ((tid & 31) == 0) { wmax[wp] = 0.f; if (wp < 5) wmax[8 + wp] = 0.f; }
                    wmax[wp] = fmaxf(wmax[wp], sqmax);
                    if (wp < 5) wmax[8 + wp] = fmaxf(wmax[8 + wp], ringmax);
#else
                    const unsigned m1 = __reduce_max_sync(0xffffffffu, __float_as_uint(sqmax));
                    if ((tid & 31) == 0) wmax[wp] = __uint_as_float(m1);
                    if (wp < 5) {
                        const unsigned m2 = __reduce_max_sync(0xffffffffu, __float_as_uint(ringmax));
                        if ((tid & 31) == 0) wmax[8 + wp] = __uint_as_float(m2);
                    }
#endif
                }
            }
        } else {
            SS_THREADS(tid)
            {
                const int tx = tid & 15, ty = tid >> 4;
                for (int q = 0; q < 16; ++q) {
                    const int qq = 4 * ty + (q >> 2) + 2, pp = 4 * tx + (q & 3) + 2;
                    const int y = y0 - 2 + qq, x = x0 - 2 + pp;
                    if (y >= p.h || x >= p.w) continue;
                    float lap[3];
                    SQ[qq * SS_SQS + pp] = ss_lap_px_generic(P0, P1, p.k25, qq, pp, y, x, i0, j0, x >= wz_vec, lap);
                    for (int ch = 0; ch < 3; ++ch)
                        P0[ch * SS_P0P + (qq + SS_P0Y - 2) * SS_P0S + pp + SS_P0X - 2] = lap[ch];
                }
                for (int i = tid; i < 4 * (SS_SQW + SS_SQH - 4); i += SS_NT) {
                    int qq, pp;
                    if (i < 2 * SS_SQW) { qq = i / SS_SQW; pp = i - qq * SS_SQW; }
                    else if (i < 4 * SS_SQW) { const int k = i - 2 * SS_SQW; qq = SS_SQH - 2 + k / SS_SQW; pp = k % SS_SQW; }
                    else { const int k = i - 4 * SS_SQW; qq = 2 + (k >> 2); pp = (k & 3) < 2 ? (k & 3) : SS_SQW - 4 + (k & 3); }
                    const int y = y0 - 2 + qq, x = x0 - 2 + pp;
                    if (y < 0 || y >= p.h || x < 0 || x >= p.w) continue;
                    float lap[3];
                    SQ[qq * SS_SQS + pp] = ss_lap_px_generic(P0, P1, p.k25, qq, pp, y, x, i0, j0, x >= wz_vec, lap);
                }
            }
        }
        SS_SYNC();
        if (!interior) {
            // out-of-image gray^2 positions: REFLECT101 of the energy filter at the size of g_l
            SS_THREADS(tid)
            {
                for (int i = tid; i < SS_SQH * SS_SQW; i += SS_NT) {
                    const int qq = i / SS_SQW, pp = i - qq * SS_SQW;
                    const int y = y0 - 2 + qq, x = x0 - 2 + pp;
                    if (y >= 0 && y < p.h && x >= 0 && x < p.w) continue;
                    const int ys = y < 0 ? -y : (y >= p.h ? 2 * p.h - 2 - y : y);
                    const int xs = x < 0 ? -x : (x >= p.w ? 2 * p.w - 2 - x : x);
                    if (ys < 0 || ys >= p.h || xs < 0 || xs >= p.w) continue;
                    const int qs = ys - (y0 - 2), ps = xs - (x0 - 2);
                    if (qs < 0 || qs >= SS_SQH || ps < 0 || ps >= SS_SQW) continue;
                    SQ[qq * SS_SQS + pp] = SQ[qs * SS_SQS + ps];
                }
            }
            SS_SYNC();
        }
        // ---- S3: energy + running arg-max ----------------------------------------------------------
#ifdef SS_HOST_EMU
        if (s0_owner && f + 1 < f_end) {   // stands in for the TMA copy that runs during the energy stage
            SS_THREADS(tid) { ss_s0_async_u8(p, frame + p.in_frame_stride, x0, y0, tid, (char *)P1); }
        }
#endif
#ifdef SS_HOST_EMU
        float warp_bemin[SS_NT / 32];
        for (int wq = 0; wq < SS_NT / 32; ++wq) {
            warp_bemin[wq] = bemin[32 * wq][0];
            for (int l = 1; l < 32; ++l) warp_bemin[wq] = fminf(warp_bemin[wq], bemin[32 * wq + l][0]);
        }
#endif
        SS_THREADS(tid)
        {
            const int tx = tid & 15, ty = tid >> 4;
            const int gidx = p.frame_index0 + f;
            const bool take_all = first_frame;
            if (pref32 && f + 1 < f_end)   // P0 is not read any more in this frame
                ss_s0_bulk_f32(p, frame + p.in_frame_stride, x0, y0, tid, P0, mbar);
#ifndef SS_HOST_EMU
            if (s0_owner && f + 1 < f_end && tid == 0) {   // P1 is dead until the next S1: raw bytes of the next frame
                ss_tma_load_2d(P1, tmap, 8 + (x0 - SS_P0X) * 3, p.tma_row0 + (f + 1) * p.h + y0 - SS_P0Y, mbar, SS_P0H * SS_RAWS);
                if (f + 2 < f_end)                         // ... and the one after it as far as L2
                    ss_tma_prefetch_2d(tmap, 8 + (x0 - SS_P0X) * 3, p.tma_row0 + (f + 2) * p.h + y0 - SS_P0Y);
            }
#endif
            bool skip = false;
            if (skip_on && !first_frame) {
                // largest gray^2 the windows of this warp can contain: own rows and the two neighbouring warps' rows
                // (the window reaches 2 rows up and down) plus the whole ring
                const int wp = tid >> 5;
                float mx = fmaxf(fmaxf(wmax[8], wmax[9]), fmaxf(fmaxf(wmax[10], wmax[11]), wmax[12]));
                mx = fmaxf(mx, wmax[wp]);
                if (wp > 0) mx = fmaxf(mx, wmax[wp - 1]);
                if (wp < 7) mx = fmaxf(mx, wmax[wp + 1]);
#ifdef SS_HOST_EMU
                const float lo = warp_bemin[wp];
#else
                const float lo = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(SS_PT(bemin)[0])));
#endif
                skip = ss_mul(mx, p.skip_factor) < lo;
            }
            float e[16];
            if (skip) {
                // no pixel of this warp can win: nothing to do
#ifdef SS_HOST_EMU
                ss_emu_skipped_threads++;
#endif
            } else {
            if (fast) {
#pragma unroll
                for (int q = 0; q < 16; ++q) e[q] = 0.f;
                const float *src = SQ + (4 * ty) * SS_SQS + 4 * tx;
                float vv[2][8];
                {
                    const ss_f4 a = *(const ss_f4 *)src, b = *(const ss_f4 *)(src + 4);
                    vv[0][0] = a.x; vv[0][1] = a.y; vv[0][2] = a.z; vv[0][3] = a.w;
                    vv[0][4] = b.x; vv[0][5] = b.y; vv[0][6] = b.z; vv[0][7] = b.w;
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (r + 1 < 8) {
                        const ss_f4 a = *(const ss_f4 *)(src + (r + 1) * SS_SQS), b = *(const ss_f4 *)(src + (r + 1) * SS_SQS + 4);
                        float *n = vv[(r + 1) & 1];
                        n[0] = a.x; n[1] = a.y; n[2] = a.z; n[3] = a.w; n[4] = b.x; n[5] = b.y; n[6] = b.z; n[7] = b.w;
                    }
                    const float *v = vv[r & 1];
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const int dy = r - o;
                        if (dy < 0 || dy > 4) continue;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
#pragma unroll
                            for (int dx = 0; dx < 5; ++dx)
                                e[o * 4 + c] = ss_fma(v[c + dx], p.k25[dy * 5 + dx], e[o * 4 + c]);
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int yy = 4 * ty + (q >> 2), xx = 4 * tx + (q & 3);
                    e[q] = ss_conv25_generic(SQ + yy * SS_SQS + xx, SS_SQS, p.k25, x0 + xx >= w_vec);
                }
            }
            // branch-free compare/update (one FSETP + FSEL + predicated OR per pixel)
            unsigned won = 0;
            if (full) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const bool w = take_all | (e[q] > SS_PT(be)[q]);
                    SS_PT(be)[q] = w ? e[q] : SS_PT(be)[q];
                    won |= w ? (1u << q) : 0u;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const bool inb = (x0 + 4 * tx + (q & 3) < p.w) & (y0 + 4 * ty + (q >> 2) < p.h);
                    const bool w = inb & (take_all | (e[q] > SS_PT(be)[q]));
                    SS_PT(be)[q] = w ? e[q] : SS_PT(be)[q];
                    won |= w ? (1u << q) : 0u;
                }
            }
            if (fast) {
                // register-blocked tiles: the parked Laplacians come back from tensor memory, one plane at a time;
                // the loads are warp-wide, so they run whenever any lane of the warp has a winner
#ifdef SS_HOST_EMU
                const bool anywon = won != 0;
#else
                const bool anywon = __any_sync(0xffffffffu, won != 0);
#endif
                if (anywon) {
                    const long long o0 = (long long)(y0 + 4 * ty) * p.w + x0 + 4 * tx;
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (won & (1u << q)) st_idx[o0 + (long long)(q >> 2) * p.w + (q & 3)] = gidx;
                    if (!(pref32 && first_frame)) {   // (else the coefficients were stored by S2)
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            float l[16];
#ifdef SS_HOST_EMU
                            for (int q = 0; q < 16; ++q) l[q] = SS_PT(tmv)[ch * 16 + q];
#else
                            ss_tm_ld16(tm_addr + ch * 16, l);
#endif
#pragma unroll
                            for (int q = 0; q < 16; ++q)
                                if (won & (1u << q))
                                    st_fused[ch * plane_px + o0 + (long long)(q >> 2) * p.w + (q & 3)] = ss_add(0.f, l[q]);
                        }
                    }
                }
            } else if (won) {
                const long long o0 = (long long)(y0 + 4 * ty) * p.w + x0 + 4 * tx;
                const float *lsrc = P0 + (4 * ty + SS_P0Y) * SS_P0S + 4 * tx + SS_P0X;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    if (won & (1u << q)) {
                        const long long o = o0 + (long long)(q >> 2) * p.w + (q & 3);
                        st_idx[o] = gidx;
                        const float *ls = lsrc + (q >> 2) * SS_P0S + (q & 3);
                        st_fused[o] = ss_add(0.f, ls[0]);
                        st_fused[plane_px + o] = ss_add(0.f, ls[SS_P0P]);
                        st_fused[2 * plane_px + o] = ss_add(0.f, ls[2 * SS_P0P]);
                    }
                }
            }
            if (won) {
                if (skip_on) {
                    float m = SS_PT(be)[0];
#pragma unroll
                    for (int q = 1; q < 16; ++q) m = fminf(m, SS_PT(be)[q]);
                    SS_PT(bemin)[0] = m;
                }
            }
            }   // !skip
            // uint8 interior tiles: S0 of the next frame.  This thread is done with the Laplacians parked in its own
            // block, and nobody reads the frame around the interior after S2, so no barrier is needed first; the raw
            // bytes of the next frame must have landed
            if (s0_owner && f + 1 < f_end) {
#ifndef SS_HOST_EMU
                ss_mbar_wait(mbar, mbar_phase);
#endif
                ss_s0_owner_u8(P0, tid, (const char *)P1);
            }
        }
        // float32 levels: the bulk copies of the next frame's patch must have landed
#ifndef SS_HOST_EMU
        if (s0_owner && f + 1 < f_end) mbar_phase ^= 1;
        if (pref32 && f + 1 < f_end) {
            ss_mbar_wait(mbar, mbar_phase);
            mbar_phase ^= 1;
        }
#endif
        SS_SYNC();
    }

    SS_THREADS(tid)
    {
        const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int x = x0 + 4 * tx + c, y = y0 + 4 * ty + r;
                if (x < p.w && y < p.h) st_e[(long long)y * p.w + x] = SS_PT(be)[r * 4 + c];
            }
    }
#ifndef SS_HOST_EMU
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "n"(SS_TM_COLS) : "memory");
#endif
}
