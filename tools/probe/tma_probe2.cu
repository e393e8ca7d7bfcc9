// Second TMA probe: vary the kernel-side protocol around a 3-D fp64 box load.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
int g_c0 = 1;
constexpr int BX = 64, BY = 16, BZ = 1, CELLS = BX * BY * BZ;

template <int V>
__global__ void kern(const __grid_constant__ CUtensorMap map, double* out, int c0, int c1, int c2) {
    extern __shared__ __align__(1024) unsigned char raw[];
    double* tile = reinterpret_cast<double*>(raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(raw + CELLS * 8);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(bar)) : "memory");
        if (V & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (V & 2) {   // copy first, then arrive (guide order)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, { %3, %4, %5 }], [%2];"
                         ::"r"(saddr(tile)), "l"(&map), "r"(saddr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(CELLS * 8) : "memory");
        } else {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(CELLS * 8) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, { %3, %4, %5 }], [%2];"
                         ::"r"(saddr(tile)), "l"(&map), "r"(saddr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
        }
    }
    if (V & 4) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(saddr(bar)) : "memory");
    } else {
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(saddr(bar)) : "memory");
    }
    for (int n = threadIdx.x; n < CELLS; n += blockDim.x) out[n] = tile[n];
}

template <int V>
int run(const CUtensorMap& map, double* out, size_t smem) {
    CK(cudaFuncSetAttribute(kern<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
    kern<V><<<1, 128, smem>>>(map, out, g_c0, 1, 3);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

int main(int argc, char** argv) {
    int v = argc > 1 ? atoi(argv[1]) : 0; if (argc > 2) g_c0 = atoi(argv[2]);
    const int X = 70, Y = 70, Z = 66;
    double* h = (double*)malloc(sizeof(double) * X * Y * Z);
    for (int n = 0; n < X * Y * Z; ++n) h[n] = n;
    double *d, *out;
    CK(cudaMalloc(&d, sizeof(double) * X * Y * Z));
    CK(cudaMalloc(&out, sizeof(double) * CELLS));
    CK(cudaMemcpy(d, h, sizeof(double) * X * Y * Z, cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult st;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st));
    auto encode = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z}; cuuint64_t strides[2] = {X * 8ull, X * Y * 8ull};
    cuuint32_t box[3] = {BX, BY, BZ}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode rc=%d\n", (int)r); return 1; }
    size_t smem = CELLS * 8 + 64;
    int rc = 1;
    switch (v) {
        case 0: rc = run<0>(map, out, smem); break;
        case 1: rc = run<1>(map, out, smem); break;
        case 2: rc = run<2>(map, out, smem); break;
        case 3: rc = run<3>(map, out, smem); break;
        case 4: rc = run<4>(map, out, smem); break;
        case 5: rc = run<5>(map, out, smem); break;
        case 7: rc = run<7>(map, out, smem); break;
    }
    if (rc) return rc;
    double* back = (double*)malloc(sizeof(double) * CELLS);
    CK(cudaMemcpy(back, out, sizeof(double) * CELLS, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int k = 0; k < BZ; ++k) for (int j = 0; j < BY; ++j) for (int i = 0; i < BX; ++i)
        if (back[i + j * BX + k * BX * BY] != (g_c0 + i) + (1 + j) * (double)X + (3 + k) * (double)X * Y) ++bad;
    printf("variant %d: launch ok, %d mismatches of %d\n", v, bad, CELLS);
    return 0;
}
