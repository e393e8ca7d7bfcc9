// Stand-alone TMA probe: which tensor-map flavours of a haloed fp64 box load work on this GPU?
// usage: tma_probe <mode>   (one mode per process: a faulting mode kills the context)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Wrapped { CUtensorMap map; double* out; };

template <int RANK>
__device__ void body(const CUtensorMap* map, double* out, int box_cells, int c0, int c1, int c2) {
    extern __shared__ __align__(128) unsigned char raw[];
    double* tile = reinterpret_cast<double*>(raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(raw + box_cells * 8);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (RANK == 7) {
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(bar)) : "memory");
    } else if (RANK == 8) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(box_cells * 8) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(saddr(tile)), "l"(out + 4096), "r"(box_cells * 8), "r"(saddr(bar)) : "memory");
        }
    } else if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(box_cells * 8) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, { %3, %4, %5 }], [%2];"
                         ::"r"(saddr(tile)), "l"(map), "r"(saddr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, { %3, %4 }], [%2];"
                         ::"r"(saddr(tile)), "l"(map), "r"(saddr(bar)), "r"(c0), "r"(c1) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(saddr(bar)) : "memory");
    for (int n = threadIdx.x; n < box_cells; n += blockDim.x) out[n] = tile[n];
}

__global__ void k_direct3(const __grid_constant__ CUtensorMap map, double* out, int cells, int c0, int c1, int c2) { body<3>(&map, out, cells, c0, c1, c2); }
__global__ void k_direct2(const __grid_constant__ CUtensorMap map, double* out, int cells, int c0, int c1, int c2) { body<2>(&map, out, cells, c0, c1, c2); }
__global__ void k_mbar(const __grid_constant__ CUtensorMap map, double* out, int cells, int c0, int c1, int c2) { body<7>(&map, out, cells, c0, c1, c2); }
__global__ void k_bulk(const __grid_constant__ CUtensorMap map, double* out, int cells, int c0, int c1, int c2) { body<8>(&map, out, cells, c0, c1, c2); }
__global__ void k_global3(const CUtensorMap* map, double* out, int cells, int c0, int c1, int c2) { body<3>(map, out, cells, c0, c1, c2); }
__global__ void k_wrapped3(const __grid_constant__ Wrapped w, int cells, int c0, int c1, int c2) { body<3>(&w.map, w.out, cells, c0, c1, c2); }

int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int X = argc > 6 ? atoi(argv[6]) : 70, Y = 70, Z = 66;               // padded 64x64x60
    const int bx = argc > 2 ? atoi(argv[2]) : 68, by = argc > 3 ? atoi(argv[3]) : 20, bz = argc > 4 ? atoi(argv[4]) : 1;
    double* h = (double*)malloc(sizeof(double) * X * Y * Z);
    for (int n = 0; n < X * Y * Z; ++n) h[n] = n;
    double *d, *out;
    CK(cudaMalloc(&d, sizeof(double) * X * Y * Z));
    CK(cudaMalloc(&out, sizeof(double) * (bx * by * bz + 8192)));
    CK(cudaMemcpy(d, h, sizeof(double) * X * Y * Z, cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult st;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st));
    auto encode = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    CUtensorMap map;
    cuuint64_t dims[3] = {X, Y, Z}; cuuint64_t strides[2] = {X * 8ull, X * Y * 8ull};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz}; cuuint32_t es[3] = {1, 1, 1};
    CUtensorMapDataType type = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    if (mode == 2) type = CU_TENSOR_MAP_DATA_TYPE_INT64;
    if (mode == 3) type = CU_TENSOR_MAP_DATA_TYPE_UINT64;
    if (mode == 9) type = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    int rank = (mode == 4) ? 2 : 3;
    CUresult r = encode(&map, type, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)(argc > 5 ? atoi(argv[5]) : 3), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d encode rc=%d desc:", mode, (int)r); for (int q = 0; q < 16; ++q) printf(" %016llx", ((unsigned long long*)&map)[q]); printf("\n"); fflush(stdout);
    if (r != CUDA_SUCCESS) return 1;
    int cells = bx * by * bz;
    size_t smem = cells * 8 + 64;
    CK(cudaFuncSetAttribute(k_direct3, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
    CK(cudaFuncSetAttribute(k_direct2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
    CK(cudaFuncSetAttribute(k_wrapped3, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
    CK(cudaFuncSetAttribute(k_mbar, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
    if (mode == 10) {
        CUtensorMap* dmap; CK(cudaMalloc(&dmap, sizeof(CUtensorMap))); CK(cudaMemcpy(dmap, &map, sizeof(map), cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(k_global3, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
        k_global3<<<1, 128, smem>>>(dmap, out, cells, 1, 1, 3);
    } else
    if (mode == 7) k_mbar<<<1, 128, smem>>>(map, out, cells, 1, 1, 3);
    else if (mode == 8) k_bulk<<<1, 128, smem>>>(map, out, cells, 1, 1, 3);
    else if (mode == 4) k_direct2<<<1, 128, smem>>>(map, out, cells, 1, 1, 3);
    else if (mode == 6) { Wrapped w; w.map = map; w.out = out; k_wrapped3<<<1, 128, smem>>>(w, cells, 1, 1, 3); }
    else k_direct3<<<1, 128, smem>>>(map, out, cells, 1, 1, 3);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    double* back = (double*)malloc(sizeof(double) * cells);
    CK(cudaMemcpy(back, out, sizeof(double) * cells, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int k = 0; k < bz; ++k) for (int j = 0; j < by; ++j) for (int i = 0; i < bx; ++i) {
        double want = (1 + i) + (1 + j) * (double)X + (3 + k) * (double)X * Y;
        if (back[i + j * bx + k * bx * by] != want) ++bad;
    }
    printf("mode %d OK launch, %d mismatches of %d\n", mode, bad, cells);
    return 0;
}
