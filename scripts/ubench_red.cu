// micro-benchmark: throughput of coalesced, distinct-address vector REDs (REDG.ADD.F32x2) vs plain stores
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_red(float2* g, long long n, int reps) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; r++)
    for (long long q = i; q < n; q += stride) atomicAdd(&g[q], make_float2(1.f, 2.f));
}
__global__ void k_red4(float4* g, long long n, int reps) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; r++)
    for (long long q = i; q < n; q += stride) atomicAdd(&g[q], make_float4(1.f, 2.f, 3.f, 4.f));
}
__global__ void k_st(float2* g, long long n, int reps) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; r++)
    for (long long q = i; q < n; q += stride) g[q] = make_float2(1.f + r, 2.f);
}
// tile-shaped: each CTA REDs a 22^3 padded tile into a 256^3 grid (like the spreader flush)
__global__ void k_red_tile(float2* g, int P, int bs, int Nt) {
  int tile = blockIdx.x; int nb = Nt / bs;
  int tx = tile % nb, ty = (tile / nb) % nb, tz = tile / (nb * nb);
  int x0 = tx * bs - (P - bs) / 2, y0 = ty * bs - (P - bs) / 2, z0 = tz * bs - (P - bs) / 2;
  for (int q = threadIdx.x; q < P * P * P; q += blockDim.x) {
    int x = q % P, r = q / P, y = r % P, z = r / P;
    int gx = (x0 + x + Nt) % Nt, gy = (y0 + y + Nt) % Nt, gz = (z0 + z + Nt) % Nt;
    atomicAdd(&g[((long long)gz * Nt + gy) * Nt + gx], make_float2(1.f, 2.f));
  }
}
int main() {
  long long n = 1ll << 24;
  float2* g; cudaMalloc(&g, n * 8); cudaMemset(g, 0, n * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
  for (int it = 0; it < 2; it++) {
    cudaEventRecord(a); k_red<<<148 * 8, 256>>>(g, n, 4); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("RED.F32x2 coalesced: %.1f us per 2^24 cells -> %.1f Gred/s, %.0f GB/s payload\n", ms * 250, n * 4 / ms / 1e6, n * 4 * 8 / ms / 1e6);
    cudaEventRecord(a); k_red4<<<148 * 8, 256>>>((float4*)g, n / 2, 4); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("RED.F32x4 coalesced: %.1f us per 2^24 cells -> %.0f GB/s payload\n", ms * 250, n * 4 * 8 / ms / 1e6);
    cudaEventRecord(a); k_st<<<148 * 8, 256>>>(g, n, 4); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("ST.64 coalesced:     %.1f us per 2^24 cells -> %.0f GB/s\n", ms * 250, n * 4 * 8 / ms / 1e6);
    cudaEventRecord(a); k_red_tile<<<4096, 256>>>(g, 22, 16, 256); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("tile flush 22^3 x 4096 tiles (43.6M REDs): %.1f us\n", ms * 1000);
    cudaEventRecord(a); k_red_tile<<<32768, 128>>>(g, 14, 8, 256); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("tile flush 14^3 x 32768 tiles (89.9M REDs): %.1f us\n", ms * 1000);
  }
  return 0;
}
