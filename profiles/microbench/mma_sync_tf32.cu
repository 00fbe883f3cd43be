// Microbenchmark (measurement tooling, not product code): peak rate of the LEGACY warp-level tensor path
// (mma.sync.aligned.m16n8k8 tf32, SASS HMMA) on sm_100a.  Question it answers for the attention kernels: for
// (slate, head) problems as small as S = 240, dk = 32 the tcgen05 kernels are bound by mbarrier / TMEM hand-off
// latency (12 % tensor-pipe active), so a register-resident FlashAttention-2-style kernel on mma.sync could win IF the
// legacy path still delivers a useful fraction of the tcgen05 rate.  Attention needs 120.8 GFLOP forward / 302 GFLOP
// backward per layer at B = 4096.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_tf32 mma_sync_tf32.cu && ./mma_sync_tf32
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ACCS>
__global__ void __launch_bounds__(256) mma_loop(int iters, float* out) {
  float c[ACCS][4];
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + 0.001f * float(threadIdx.x + i));
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + 0.001f * float(threadIdx.x + i));
#pragma unroll
  for (int k = 0; k < ACCS; ++k) c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ACCS; ++k) mma_tf32(c[k], a, b);       // ACCS independent accumulator chains per warp
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < ACCS; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  if (s == 12345.678f) out[0] = s;
}

template <int ACCS>
static void run(int warps_per_sm_times8) {
  const int iters = 4096;
  float* out;
  cudaMalloc(&out, 4);
  const int grid = 148 * warps_per_sm_times8;      // 8 warps per CTA
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mma_loop<ACCS><<<grid, 256>>>(16, out);
  cudaEventRecord(e0);
  mma_loop<ACCS><<<grid, 256>>>(iters, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 2.0 * 16 * 8 * 8 * double(ACCS) * iters * 8.0 * grid;
  printf("accumulators/warp %d, %d warps/SM: %.3f ms, %.1f TFLOP/s tf32 (mma.sync)\n", ACCS, 8 * warps_per_sm_times8, ms,
         flop / (ms * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  run<4>(1); run<8>(1); run<8>(2); run<16>(2); run<8>(4); run<8>(8);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
