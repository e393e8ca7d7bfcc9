// The CUDA programming guide's TMA example (libcu++ wrappers), as an environment sanity check.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int SW = 32, SH = 8;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out) {
    __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int n = threadIdx.x; n < SW * SH; n += blockDim.x) out[n] = smem_buffer[n / SW][n % SW];
}
int main() {
    const int W = 256, H = 64;
    int* h = new int[W * H];
    for (int n = 0; n < W * H; ++n) h[n] = n;
    int *d, *out;
    CK(cudaMalloc(&d, W * H * 4)); CK(cudaMalloc(&out, SW * SH * 4));
    CK(cudaMemcpy(d, h, W * H * 4, cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult st;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st));
    auto encode = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    CUtensorMap map;
    cuuint64_t dims[2] = {W, H}; cuuint64_t strides[1] = {W * 4};
    cuuint32_t box[2] = {SW, SH}; cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    kernel<<<1, 128>>>(map, 16, 4, out);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    int back[SW * SH];
    CK(cudaMemcpy(back, out, sizeof(back), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int j = 0; j < SH; ++j) for (int i = 0; i < SW; ++i) if (back[j * SW + i] != (4 + j) * W + 16 + i) ++bad;
    printf("guide example: %d mismatches\n", bad);
    return 0;
}
