// standalone check of the TMA path used by ss_level_kernel<uint8_t>.  Finding (B200): the byte column of a box
// origin must be a multiple of 16 -- c0 = 168 raises "illegal instruction", c0 = 0 / 176 work; hence the frame
// store keeps every row 8 bytes into its pitch (8 + 3*(x0-8) = 16*k for x0 = 64*j).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
#define SS_NT 256
#include "../shinestacker_b200/csrc/ss_kernels.cuh"

__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, unsigned char *out, int variant)
{
    extern __shared__ __align__(128) float sm[];
    uint64_t *mbar = (uint64_t *)(sm + 8192);
    if (threadIdx.x == 0) ss_mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) ss_tma_load_2d(sm, &tmap, c0, c1, mbar, 76 * 240);
    ss_mbar_wait(mbar, 0);
    for (int i = threadIdx.x; i < 76 * 240; i += blockDim.x) out[i] = ((unsigned char *)sm)[i];
}

int main()
{
    const int pitch = 1152, rows = 1536;
    std::vector<unsigned char> h(pitch * rows);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 7 + i / pitch) & 0xff);
    unsigned char *d, *o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, 76 * 240);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry %p q=%d\n", fn, (int)q);
    alignas(64) CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {240, 76}, es[2] = {1, 1};
    CUresult r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    for (int c0 : {0, 176, 368}) {
        k<<<1, 256, 40000>>>(tm, c0, 58, o, 0);
        cudaError_t e = cudaDeviceSynchronize();
        printf("c0=%d launch: %s\n", c0, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<unsigned char> res(76 * 240);
        cudaMemcpy(res.data(), o, res.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < 76; ++rr)
            for (int cc = 0; cc < 240; ++cc)
                if (res[rr * 240 + cc] != h[(size_t)(58 + rr) * pitch + c0 + cc]) ++bad;
        printf("c0=%d mismatches %d\n", c0, bad);
    }
    return 0;
}
