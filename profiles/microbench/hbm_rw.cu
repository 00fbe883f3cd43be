// Microbenchmark (measurement tooling, not product code): pure-read, pure-write and copy bandwidth of HBM, to check
// the asymmetric bound  t = read_bytes / R + write_bytes / W  that the per-launch ncu numbers suggest
// (profiles/README.md: R ~ 6.9 TB/s, W ~ 3.25 TB/s, copy peak 6.49 TB/s = 2 x 3.25).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_rw hbm_rw.cu && ./hbm_rw
#include <cstdio>
#include <cuda_runtime.h>

__global__ void read_kernel(const float4* __restrict__ p, size_t n, float* sink) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float4 v = p[i];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) *sink = acc.x;   // never true: keeps the loads alive
}
__global__ void write_kernel(float4* __restrict__ p, size_t n) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) p[i] = v;
}
__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) b[i] = a[i];
}
// 1 read : k writes (k = 3 models the QKV projection, k = 4 the first FFN linear)
__global__ void fanout_kernel(const float4* __restrict__ a, float4* __restrict__ b, size_t n, int k) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float4 v = a[i];
    for (int j = 0; j < k; ++j) b[size_t(j) * n + i] = v;
  }
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const size_t bytes = size_t(2) << 30;             // 2 GiB per buffer: far above the 126 MB L2
  const size_t n = bytes / sizeof(float4);
  float4 *a, *b;
  float* sink;
  if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, 4 * bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&sink, 4);
  cudaMemset(a, 0, bytes);
  cudaMemset(b, 0, 4 * bytes);
  const int grid = 148 * 16, block = 512, reps = 10;
  const double gb = double(bytes) / 1e9;
  float t;
  t = time_ms([&] { read_kernel<<<grid, block>>>(a, n, sink); }, reps);
  printf("read   %.3f ms  %.0f GB/s\n", t, gb / (t * 1e-3));
  t = time_ms([&] { write_kernel<<<grid, block>>>(b, n); }, reps);
  printf("write  %.3f ms  %.0f GB/s\n", t, gb / (t * 1e-3));
  t = time_ms([&] { copy_kernel<<<grid, block>>>(a, b, n); }, reps);
  printf("copy   %.3f ms  %.0f GB/s (read+write)\n", t, 2 * gb / (t * 1e-3));
  for (int k = 3; k <= 4; ++k) {
    t = time_ms([&] { fanout_kernel<<<grid, block>>>(a, b, n, k); }, reps);
    printf("1r:%dw  %.3f ms  %.0f GB/s total, model read/6.9+write/3.25 = %.3f ms\n", k, t, (1 + k) * gb / (t * 1e-3),
           (gb / 6900.0 + k * gb / 3250.0) * 1e3);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
