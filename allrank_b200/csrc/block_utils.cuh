// Block/warp-level primitives shared by the slate kernels (losses, metrics).
// One CTA owns one slate; all helpers assume blockDim.x is a multiple of 32 and <= 1024.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace arb {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// Block reductions through a 32-slot scratch array in shared memory.  Every thread gets the result.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  T r = (lane < nw) ? scratch[lane] : T(0);
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : -CUDART_INF_F;
  r = warp_max(r);
  return r;
}

// Order-preserving map float -> uint32 (ascending).  -0.0 is canonicalised to +0.0 so that it ties
// with +0.0 exactly like a float comparison does.
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// 64-bit key: descending by score, ties by ascending position (a stable descending sort).
__device__ __forceinline__ uint64_t desc_key(float score, uint32_t pos) {
  uint32_t k = isnan(score) ? 0u : ~float_to_ordered(score);  // NaN ranks first, like torch.sort(descending=True)
  return (uint64_t(k) << 32) | pos;
}

// In-place ascending bitonic sort of n (power of two) keys in shared memory by the whole block.
template <typename K>
__device__ __forceinline__ void bitonic_sort(K* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const bool up = ((i & k) == 0);
          const K a = keys[i], b = keys[p];
          if ((a > b) == up) { keys[i] = b; keys[p] = a; }
        }
      }
      __syncthreads();
    }
  }
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// 2^x - 1 the way the reference's gain function evaluates it; exact for the integer labels of LTR data.
__device__ __forceinline__ float pow2_minus_1(float x) {
  const float r = rintf(x);
  if (r == x && x >= 0.0f && x <= 30.0f) return float(1u << int(x)) - 1.0f;
  return exp2f(x) - 1.0f;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace arb
