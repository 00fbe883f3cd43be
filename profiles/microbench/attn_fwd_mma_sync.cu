// Prototype + microbenchmark (measurement tooling, not product code): the attention forward of one (slate, head) as a
// register-resident FlashAttention-2-style kernel on the LEGACY warp-level tensor path (mma.sync.m16n8k8 tf32).
//
// Why: the product kernel (allrank_b200/csrc/attention_fused.cu, tcgen05 + TMEM) takes 1.30 ms per launch at
// B = 4096 (16 384 (slate, head) problems of S = 240, dk = 32) = 0.29 of its HBM bound with the tensor pipe 12 % active:
// it is bound by the TMA -> MMA -> tcgen05.ld -> softmax -> tcgen05.st -> MMA hand-off chain of each CTA, not by
// throughput.  Here a warp owns 16 query rows and keeps their whole score row (240 keys = 30 accumulator tiles) in
// registers: no TMEM round trips, no mbarriers, K / V staged once per CTA in shared memory (tf32-rounded, padded rows:
// all fragment loads are bank-conflict-free), P feeds the second product straight from the accumulator registers by
// relabelling the contraction index (accumulator column 2t, 2t+1 <-> A-fragment column t, t+4; V rows loaded with the
// same permutation).  Work: 3 600 MMAs per problem; at 512 FMA/clk/SM the launch needs >= 0.40 ms of tensor time.
//
// The program checks the kernel against a plain fp32 reference on a small case, then times the B = 4096 shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o attn_fwd_mma_sync attn_fwd_mma_sync.cu && ./attn_fwd_mma_sync
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int S = 240;        // slate length (multiple of 16)
constexpr int DK = 32;        // head width
constexpr int NT = S / 8;     // 8-key accumulator tiles per score row
constexpr int ND = DK / 8;    // 8-column accumulator tiles of the output
constexpr int KS = DK / 8;    // k-steps of Q K^T
constexpr int PITCH = DK + 4; // shared-memory row pitch in floats: (4 g + t) and (8 t + g) bank patterns are conflict-free
constexpr int WARPS = 4;

__device__ __forceinline__ uint32_t tf32(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                    uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// qkv: [B*S, 3*d_model] packed (Q | K | V, heads side by side), mask [B, S] (1 = padded key), ctx: [B*S, d_model].
__global__ void __launch_bounds__(WARPS * 32) attn_fwd_mma(const float* __restrict__ qkv, const uint8_t* __restrict__ mask,
                                                          float* __restrict__ ctx, float* __restrict__ stat_max,
                                                          float* __restrict__ stat_sum, int n_heads, int d_model,
                                                          float scale_log2e) {
  extern __shared__ uint32_t smem[];
  uint32_t* Ks = smem;                       // [S][PITCH] tf32 bits
  uint32_t* Vs = smem + S * PITCH;           // [S][PITCH]
  uint32_t* valid = Vs + S * PITCH;          // [NT] one byte-sized bit set per 8-key tile (bit j = key 8*nt + j is real)
  const int head = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t pitch = size_t(3) * d_model;
  const float* base = qkv + size_t(b) * S * pitch + size_t(head) * DK;

  for (int i = tid; i < S * (DK / 4); i += WARPS * 32) {     // K and V rows: coalesced 128-bit loads, rounded once
    const int r = i / (DK / 4), c4 = i % (DK / 4);
    const float4 k = *reinterpret_cast<const float4*>(base + size_t(r) * pitch + d_model + 4 * c4);
    const float4 v = *reinterpret_cast<const float4*>(base + size_t(r) * pitch + 2 * d_model + 4 * c4);
    uint32_t* kd = Ks + r * PITCH + 4 * c4;
    uint32_t* vd = Vs + r * PITCH + 4 * c4;
    kd[0] = tf32(k.x); kd[1] = tf32(k.y); kd[2] = tf32(k.z); kd[3] = tf32(k.w);
    vd[0] = tf32(v.x); vd[1] = tf32(v.y); vd[2] = tf32(v.z); vd[3] = tf32(v.w);
  }
  for (int nt = tid; nt < NT; nt += WARPS * 32) {
    uint32_t bits = 0;
    for (int j = 0; j < 8; ++j) bits |= (mask[size_t(b) * S + 8 * nt + j] == 0 ? 1u : 0u) << j;
    valid[nt] = bits;
  }
  __syncthreads();

  for (int tile = warp; tile < S / 16; tile += WARPS) {
    const int r0 = 16 * tile + g, r1 = r0 + 8;
    // ---- Q fragments (row-major A: a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4))
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      qa[ks][0] = tf32(base[size_t(r0) * pitch + 8 * ks + t]);
      qa[ks][1] = tf32(base[size_t(r1) * pitch + 8 * ks + t]);
      qa[ks][2] = tf32(base[size_t(r0) * pitch + 8 * ks + t + 4]);
      qa[ks][3] = tf32(base[size_t(r1) * pitch + 8 * ks + t + 4]);
    }
    // ---- scores: acc[nt] = Q K^T over the 8 keys of tile nt (B col-major: b0 (k = t, n = g)  b1 (k = t+4, n = g))
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      const uint32_t* krow = Ks + (8 * nt + g) * PITCH + t;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) mma(acc[nt], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], krow[8 * ks], krow[8 * ks + 4]);
    }
    // ---- key mask, row maximum (c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1); a row lives in one quad)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const uint32_t bits = valid[nt] >> (2 * t);
      const bool v0 = bits & 1u, v1 = bits & 2u;
      acc[nt][0] = v0 ? acc[nt][0] * scale_log2e : -INFINITY;
      acc[nt][1] = v1 ? acc[nt][1] * scale_log2e : -INFINITY;
      acc[nt][2] = v0 ? acc[nt][2] * scale_log2e : -INFINITY;
      acc[nt][3] = v1 ? acc[nt][3] * scale_log2e : -INFINITY;
      m0 = fmaxf(m0, fmaxf(acc[nt][0], acc[nt][1]));
      m1 = fmaxf(m1, fmaxf(acc[nt][2], acc[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    // ---- exponentials, row sums, O = P V with the contraction index relabelled: A-fragment column t <-> key 2t,
    //      column t+4 <-> key 2t+1 of the 8-key group, so the accumulator registers ARE the A fragment
    float s0 = 0.f, s1 = 0.f;
    float o[ND][4], o2[ND][4];      // two accumulator sets (even / odd key groups): 8 independent MMA chains, not 4
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
      o2[nd][0] = o2[nd][1] = o2[nd][2] = o2[nd][3] = 0.f;
    }
#pragma unroll
    for (int kb = 0; kb < NT; ++kb) {
      const float e0 = ex2(acc[kb][0] - m0), e1 = ex2(acc[kb][1] - m0);
      const float e2 = ex2(acc[kb][2] - m1), e3 = ex2(acc[kb][3] - m1);
      s0 += e0 + e1;
      s1 += e2 + e3;
      const uint32_t a0 = tf32(e0), a1 = tf32(e2), a2 = tf32(e1), a3 = tf32(e3);
      const uint32_t* v0 = Vs + (8 * kb + 2 * t) * PITCH + g;      // key 2t   -> k = t
      const uint32_t* v1 = v0 + PITCH;                             // key 2t+1 -> k = t+4
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        if (kb & 1) mma(o2[nd], a0, a1, a2, a3, v0[8 * nd], v1[8 * nd]);
        else mma(o[nd], a0, a1, a2, a3, v0[8 * nd], v1[8 * nd]);
      }
    }
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      o[nd][0] += o2[nd][0]; o[nd][1] += o2[nd][1]; o[nd][2] += o2[nd][2]; o[nd][3] += o2[nd][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float i0 = 1.0f / s0, i1 = 1.0f / s1;
    float* out0 = ctx + (size_t(b) * S + r0) * d_model + size_t(head) * DK + 2 * t;
    float* out1 = ctx + (size_t(b) * S + r1) * d_model + size_t(head) * DK + 2 * t;
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      *reinterpret_cast<float2*>(out0 + 8 * nd) = make_float2(o[nd][0] * i0, o[nd][1] * i0);
      *reinterpret_cast<float2*>(out1 + 8 * nd) = make_float2(o[nd][2] * i1, o[nd][3] * i1);
    }
    if (t == 0) {       // statistics the backward pass recomputes P from (in log2 units of the scaled scores)
      const size_t so = (size_t(b) * n_heads + head) * S;
      stat_max[so + r0] = m0; stat_sum[so + r0] = s0;
      stat_max[so + r1] = m1; stat_sum[so + r1] = s1;
    }
  }
}

// plain fp32 reference: one thread per (slate, head, query)
__global__ void attn_ref(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, float* __restrict__ ctx,
                         int B, int n_heads, int d_model, float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * n_heads * S) return;
  const int q = int(idx % S), head = int((idx / S) % n_heads), b = int(idx / ((long long)S * n_heads));
  const size_t pitch = size_t(3) * d_model;
  const float* base = qkv + size_t(b) * S * pitch + size_t(head) * DK;
  float sc[S];
  float mx = -INFINITY;
  for (int k = 0; k < S; ++k) {
    float d = 0.f;
    for (int e = 0; e < DK; ++e) d += base[size_t(q) * pitch + e] * base[size_t(k) * pitch + d_model + e];
    sc[k] = mask[size_t(b) * S + k] ? -INFINITY : d * scale;
    mx = fmaxf(mx, sc[k]);
  }
  float sum = 0.f;
  for (int k = 0; k < S; ++k) { sc[k] = expf(sc[k] - mx); sum += sc[k]; }
  for (int e = 0; e < DK; ++e) {
    float o = 0.f;
    for (int k = 0; k < S; ++k) o += sc[k] * base[size_t(k) * pitch + 2 * d_model + e];
    ctx[(size_t(b) * S + q) * d_model + size_t(head) * DK + e] = o / sum;
  }
}

static void fill(std::vector<float>& v, unsigned seed) {
  unsigned s = seed;
  for (auto& x : v) {
    s = s * 1664525u + 1013904223u;
    x = (float((s >> 8) & 0xffff) / 65535.0f - 0.5f) * 2.0f;      // uniform in [-1, 1]
  }
}

int main() {
  const int n_heads = 4, d_model = n_heads * DK;
  const float scale = 1.0f / sqrtf(float(DK));
  const size_t smem = (size_t(2) * S * PITCH + NT) * sizeof(uint32_t);
  cudaFuncSetAttribute(attn_fwd_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, attn_fwd_mma);
  int ctas = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, attn_fwd_mma, WARPS * 32, smem);
  printf("attn_fwd_mma: %d registers/thread, %zu B smem/CTA, %d CTAs/SM\n", fa.numRegs, smem, ctas);

  // ---- correctness on a small batch (ragged slates: a random tail of every slate is padding)
  {
    const int B = 6;
    std::vector<float> h(size_t(B) * S * 3 * d_model);
    fill(h, 7u);
    std::vector<uint8_t> hm(size_t(B) * S, 0);
    for (int b = 0; b < B; ++b)
      for (int k = 60 + 29 * b; k < S; ++k) hm[size_t(b) * S + k] = 1;
    float *qkv, *c0, *c1, *sm, *ss;
    uint8_t* m;
    cudaMalloc(&qkv, h.size() * 4); cudaMalloc(&c0, size_t(B) * S * d_model * 4); cudaMalloc(&c1, size_t(B) * S * d_model * 4);
    cudaMalloc(&sm, size_t(B) * n_heads * S * 4); cudaMalloc(&ss, size_t(B) * n_heads * S * 4);
    cudaMalloc(&m, hm.size());
    cudaMemcpy(qkv, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(m, hm.data(), hm.size(), cudaMemcpyHostToDevice);
    attn_fwd_mma<<<dim3(n_heads, B), WARPS * 32, smem>>>(qkv, m, c0, sm, ss, n_heads, d_model, scale * 1.4426950408889634f);
    const long long n = (long long)B * n_heads * S;
    attn_ref<<<unsigned((n + 63) / 64), 64>>>(qkv, m, c1, B, n_heads, d_model, scale);
    std::vector<float> a(size_t(B) * S * d_model), r(a.size());
    cudaMemcpy(a.data(), c0, a.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(r.data(), c1, r.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0.0;
    for (size_t i = 0; i < a.size(); ++i) worst = fmax(worst, fabs(double(a[i]) - double(r[i])));
    printf("max |mma.sync - fp32 reference| = %.3e  (%s; TF32 operands: expect ~1e-3)\n", worst,
           worst < 5e-3 ? "ok" : "MISMATCH");
    cudaFree(qkv); cudaFree(c0); cudaFree(c1); cudaFree(sm); cudaFree(ss); cudaFree(m);
    if (!(worst < 5e-3)) return 2;
  }
  // ---- timing at the bench shape: B = 4096 slates x 4 heads, mean slate length ~120 of 240 like the synthetic data
  {
    const int B = 4096;
    const size_t nq = size_t(B) * S * 3 * d_model;
    float *qkv, *c0, *sm, *ss;
    uint8_t* m;
    cudaMalloc(&qkv, nq * 4); cudaMalloc(&c0, size_t(B) * S * d_model * 4);
    cudaMalloc(&sm, size_t(B) * n_heads * S * 4); cudaMalloc(&ss, size_t(B) * n_heads * S * 4);
    cudaMalloc(&m, size_t(B) * S);
    std::vector<float> h(size_t(S) * 3 * d_model * 64);
    fill(h, 11u);
    for (size_t off = 0; off < nq; off += h.size())
      cudaMemcpy(qkv + off, h.data(), std::min(h.size(), nq - off) * 4, cudaMemcpyHostToDevice);
    std::vector<uint8_t> hm(size_t(B) * S, 0);
    for (int b = 0; b < B; ++b)
      for (int k = 40 + (b * 37) % 200; k < S; ++k) hm[size_t(b) * S + k] = 1;
    cudaMemcpy(m, hm.data(), hm.size(), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i)
      attn_fwd_mma<<<dim3(n_heads, B), WARPS * 32, smem>>>(qkv, m, c0, sm, ss, n_heads, d_model, scale * 1.4426950408889634f);
    const int reps = 10;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i)
      attn_fwd_mma<<<dim3(n_heads, B), WARPS * 32, smem>>>(qkv, m, c0, sm, ss, n_heads, d_model, scale * 1.4426950408889634f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double flop = 4.0 * S * S * DK * double(n_heads) * B;
    printf("B = %d: %.3f ms per launch (product attn_fwd2_kernel: 1.30 ms), %.1f TFLOP/s tf32, %.2f TB/s of Q/K/V/O traffic\n",
           B, ms, flop / (ms * 1e-3) / 1e12, 4.0 * double(B) * S * d_model * 4 / (ms * 1e-3) / 1e12);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
