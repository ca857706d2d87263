// FP32 issue-rate microbenchmarks for sm_100a: which FMA form reaches the pipe's peak?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ float ck[32];

#define ITERS 4096

template <int CHAINS>
__global__ void k_ffma_reg(float *out, float k0, float k1, float k2, float k3)
{
    float acc[CHAINS];
    float v[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { acc[i] = threadIdx.x * 1e-3f + i; v[i] = 1.0f + i * 1e-3f + threadIdx.x * 1e-6f; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = __fmaf_rn(v[i], k0, acc[i]);
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = __fmaf_rn(v[i], k1, acc[i]);
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = __fmaf_rn(v[i], k2, acc[i]);
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = __fmaf_rn(v[i], k3, acc[i]);
    }
    float s = 0;
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void k_ffma_const(float *out)
{
    float acc[CHAINS];
    float v[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { acc[i] = threadIdx.x * 1e-3f + i; v[i] = 1.0f + i * 1e-3f + threadIdx.x * 1e-6f; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = __fmaf_rn(v[i], ck[0], acc[i]);
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = __fmaf_rn(v[i], ck[1], acc[i]);
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = __fmaf_rn(v[i], ck[2], acc[i]);
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = __fmaf_rn(v[i], ck[3], acc[i]);
    }
    float s = 0;
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}

// CHAINS packed chains = 2*CHAINS scalar chains
template <int CHAINS>
__global__ void k_ffma2(float *out, float k0, float k1, float k2, float k3)
{
    unsigned long long acc[CHAINS], v[CHAINS];
    for (int i = 0; i < CHAINS; ++i) {
        acc[i] = pack2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
        v[i] = pack2(1.0f + i * 1e-3f, 1.0f + threadIdx.x * 1e-6f);
    }
    const unsigned long long K0 = pack2(k0, k0), K1 = pack2(k1, k1), K2 = pack2(k2, k2), K3 = pack2(k3, k3);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma2(v[i], K0, acc[i]);
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma2(v[i], K1, acc[i]);
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma2(v[i], K2, acc[i]);
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma2(v[i], K3, acc[i]);
    }
    float s = 0;
    for (int i = 0; i < CHAINS; ++i) s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA chains fed from shared memory with LDS.128: 4 loads feed 16 values; each value used U times
template <int U>
__global__ void k_ffma_lds(float *out, float k0)
{
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    float acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = i;
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        float4 a = sm[idx & 1023], b = sm[(idx + 256) & 1023];
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rn(v[(i + u) & 7], ck[u & 31], acc[i]);
        idx += 7;
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + k0;
}

template <typename F>
static void timeit(const char *name, double fma_per_thread, int blocks, int threads, F launch)
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
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double fmas = fma_per_thread * blocks * threads;
    double tf = fmas / (ms * 1e-3) / 1e12;
    printf("%-34s %8.3f ms  %7.2f TFMA/s  = %6.1f FMA/clk/SM @%d MHz (nominal)\n", name, ms, tf,
           fmas / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
}

int main()
{
    float hk[32];
    for (int i = 0; i < 32; ++i) hk[i] = 1e-7f * (i + 1);
    cudaMemcpyToSymbol(ck, hk, sizeof(hk));
    float *out; cudaMalloc(&out, 148 * 16 * 1024 * 4);
    const int B = 148 * 8, T = 256;
    timeit("FFMA reg x8 chains", 8.0 * 4 * ITERS, B, T, [&] { k_ffma_reg<8><<<B, T>>>(out, 1e-7f, 2e-7f, 3e-7f, 4e-7f); });
    timeit("FFMA reg x16 chains", 16.0 * 4 * ITERS, B, T, [&] { k_ffma_reg<16><<<B, T>>>(out, 1e-7f, 2e-7f, 3e-7f, 4e-7f); });
    timeit("FFMA const x8 chains", 8.0 * 4 * ITERS, B, T, [&] { k_ffma_const<8><<<B, T>>>(out); });
    timeit("FFMA const x16 chains", 16.0 * 4 * ITERS, B, T, [&] { k_ffma_const<16><<<B, T>>>(out); });
    timeit("FFMA2 x4 packed (8 chains)", 8.0 * 4 * ITERS, B, T, [&] { k_ffma2<4><<<B, T>>>(out, 1e-7f, 2e-7f, 3e-7f, 4e-7f); });
    timeit("FFMA2 x8 packed (16 chains)", 16.0 * 4 * ITERS, B, T, [&] { k_ffma2<8><<<B, T>>>(out, 1e-7f, 2e-7f, 3e-7f, 4e-7f); });
    timeit("FFMA const + 2xLDS.128 per 8 (U=1)", 8.0 * 1 * ITERS, B, T, [&] { k_ffma_lds<1><<<B, T>>>(out, 0.f); });
    timeit("FFMA const + 2xLDS.128 per 16 (U=2)", 8.0 * 2 * ITERS, B, T, [&] { k_ffma_lds<2><<<B, T>>>(out, 0.f); });
    timeit("FFMA const + 2xLDS.128 per 32 (U=4)", 8.0 * 4 * ITERS, B, T, [&] { k_ffma_lds<4><<<B, T>>>(out, 0.f); });
    timeit("FFMA const + 2xLDS.128 per 64 (U=8)", 8.0 * 8 * ITERS, B, T, [&] { k_ffma_lds<8><<<B, T>>>(out, 0.f); });
    cudaDeviceSynchronize();
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
