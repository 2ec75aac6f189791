// TMA probe #2: the CUDA programming guide's own pattern (cuda::barrier + cde::cp_async_bulk_tensor_2d_global_to_shared)
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int SW = 32, SH = 16;
__global__ void kern(const __grid_constant__ CUtensorMap tensor_map, int x, int y, float* out)
{
    __shared__ alignas(128) float smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SW * SH; i += blockDim.x) out[i] = smem_buffer[i / SW][i % SW];
}
int main()
{
    const int W = 128, H = 4096;
    float* g; cudaMalloc(&g, (size_t)W * H * 4);
    float* h = new float[W * H];
    for (int i = 0; i < W * H; i++) h[i] = (float)i;
    cudaMemcpy(g, h, (size_t)W * H * 4, cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, SW * SH * 4);
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[2] = {W, H}; cuuint64_t str[1] = {W * 4}; cuuint32_t box[2] = {SW, SH}; cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    kern<<<1, 128>>>(tm, 32, 16, out);
    cudaError_t e = cudaDeviceSynchronize();
    float o[2] = {0, 0};
    if (e == cudaSuccess) cudaMemcpy(o, out, 8, cudaMemcpyDeviceToHost);
    printf("guide pattern: encode=%d sync=%s first=%g expect=%g\n", (int)r, cudaGetErrorString(e), o[0], (double)(16 * W + 32));
    return 0;
}
