// Conversion-cost microbenchmark for sm_100a: how many FFMA/clk/SM survive when u8 -> f32 conversions are mixed
// into the FFMA stream, as I2F.U8 (one instruction, XU pipe) or as PRMT + FADD (ALU + FMA pipes)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o mb_cvt mb_cvt.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__constant__ float ck[32];
#define ITERS 2048
__device__ __forceinline__ float cvt_i2f(uint32_t w, int k)
{
    float a;
    asm volatile("cvt.rn.f32.u8 %0, %1;" : "=f"(a) : "r"(w >> (8 * k)));
    return a;
}
__device__ __forceinline__ float cvt_prmt(uint32_t w, int k)
{
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + k)), 8388608.0f);
}
// per iteration: NW words (4 samples each) are converted, every sample feeds FPS FFMAs into 16 chains
template <int MODE, int NW, int FPS>
__global__ void k_mix(const uint32_t *in, float *out)
{
    float acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = i + threadIdx.x * 1e-3f;
    uint32_t w[NW];
    for (int i = 0; i < NW; ++i) w[i] = in[(threadIdx.x + i * 32) & 1023];
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NW; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float v;
                if (MODE == 0) v = cvt_i2f(w[i], k);
                else if (MODE == 1) v = cvt_prmt(w[i], k);
                else v = __uint_as_float(w[i] + k);   // no conversion (baseline)
#pragma unroll
                for (int f = 0; f < FPS; ++f) {
                    const int a = (i * 4 * FPS + k * FPS + f) & 15;
                    acc[a] = __fmaf_rn(v, ck[f], acc[a]);
                }
            }
            w[i] = w[i] * 3 + it;   // keep the words live and changing (1 IMAD per 4 samples)
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
static void timeit(const char *name, double fma_per_thread, double cvt_per_thread, int blocks, int threads, F launch)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    ms /= 5;
    const double clk = 1965e6;
    printf("%-44s %8.3f ms  %6.1f FMA/clk/SM  %6.1f cvt/clk/SM\n", name, ms,
           fma_per_thread * blocks * threads / (ms * 1e-3) / 148.0 / clk,
           cvt_per_thread * blocks * threads / (ms * 1e-3) / 148.0 / clk);
}
#define RUN(MODE, NW, FPS, label) timeit(label, (double)ITERS * NW * 4 * FPS, MODE == 2 ? 0.0 : (double)ITERS * NW * 4, B, T, \
    [&] { k_mix<MODE, NW, FPS><<<B, T>>>(in, out); })
int main()
{
    float hk[32];
    for (int i = 0; i < 32; ++i) hk[i] = 1e-7f * (i + 1);
    cudaMemcpyToSymbol(ck, hk, sizeof(hk));
    float *out; cudaMalloc(&out, 148 * 16 * 1024 * 4);
    uint32_t *in; cudaMalloc(&in, 4096); cudaMemset(in, 0x35, 4096);
    const int B = 148 * 2, T = 256;   // 2 CTAs x 8 warps per SM, like the level kernel
    RUN(2, 4, 8, "no conversion, 8 FFMA per value");
    RUN(2, 4, 3, "no conversion, 3 FFMA per value");
    RUN(0, 4, 1, "I2F.U8, 1 FFMA per sample");
    RUN(0, 4, 2, "I2F.U8, 2 FFMA per sample");
    RUN(0, 4, 3, "I2F.U8, 3 FFMA per sample");
    RUN(0, 4, 4, "I2F.U8, 4 FFMA per sample");
    RUN(0, 4, 6, "I2F.U8, 6 FFMA per sample");
    RUN(0, 4, 8, "I2F.U8, 8 FFMA per sample");
    RUN(1, 4, 1, "PRMT+FADD, 1 FFMA per sample");
    RUN(1, 4, 2, "PRMT+FADD, 2 FFMA per sample");
    RUN(1, 4, 3, "PRMT+FADD, 3 FFMA per sample");
    RUN(1, 4, 4, "PRMT+FADD, 4 FFMA per sample");
    RUN(1, 4, 6, "PRMT+FADD, 6 FFMA per sample");
    RUN(1, 4, 8, "PRMT+FADD, 8 FFMA per sample");
    cudaDeviceSynchronize();
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
