// Fused self-attention core for slates of up to 256 items:  ctx = softmax(mask(Q K^T / sqrt(dk))) V
// in ONE kernel, the [S,S] score / probability tile living only in tensor memory.
//
// Reference: attention() allrank/models/transformer.py:137-156 (+ the head split / concat of
// MultiHeadedAttention.forward :193-202, which here are just TMA coordinates).
//
// One CTA = one (slate b, head, 128-query tile).  Q (128 x dk), K and V (S x dk) are staged by TMA straight
// out of the packed [B*S, 3*d_model] QKV activation (4-D tensor maps: dk, item, head, slate).
//   warp 0    : TMA producer
//   warp 1    : TMEM allocation + tcgen05.mma issue
//                 S[128 x S]  = Q K^T          (kind::tf32, A/B from shared memory, accumulator in TMEM cols 0..S)
//                 O[128 x dk] = P V            (A = P read back from TMEM, B = V as an MN-major smem operand)
//   warps 2-5 : one thread per query row (= one TMEM lane): key mask, row max, exp2, row sum, P written back to
//               TMEM in place (rounded to tf32), then the O epilogue (1/rowsum) and a TMA store into the
//               concatenated-heads layout.
// Nothing of size S^2 touches HBM: per (slate, head) the kernel reads 3*S*dk*4 bytes and writes S*dk*4 (+ 8*S
// of softmax statistics for the backward pass) -- versus ~5*S^2*4 bytes for the unfused sequence.
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "attention_fused.h"
#include "common.h"
#include "sm100_ptx.cuh"

namespace arb {

constexpr int ATT_THREADS = 192;
constexpr int ATT2_SOFTMAX = 256;              // two-pass kernel: 8 softmax warps, two per TMEM lane quadrant
constexpr int ATT2_THREADS = 64 + ATT2_SOFTMAX;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t round_tf32(float x) {
  // round-to-nearest (ties away) to tf32 as "+ half an ulp of the 10-bit mantissa, then let the tensor core ignore the
  // low 13 bits" -- one integer add instead of cvt.rna.tf32.f32, which ptxas expands to three instructions here.
  // Same result as cvt.rna for every finite value (probabilities and their gradients are finite).
  return __float_as_uint(x) + 0x1000u;
}

template <int DK>
struct AttFwdSmem {
  static constexpr int NKB = (DK + 31) / 32;            // 32-wide k-blocks of the Q K^T contraction / V slabs
  static constexpr int Q_BYTES = NKB * 128 * 128;       // [kb][128 rows][128 B]
  static constexpr int K_BYTES = NKB * 256 * 128;       // [kb][256 rows][128 B]
  static constexpr int V_BYTES = NKB * 256 * 128;       // [slab][256 key rows][128 B]
  static constexpr int O_BYTES = NKB * 128 * 128;       // staging [slab][128 rows][128 B]
  static constexpr int total() { return Q_BYTES + K_BYTES + V_BYTES + O_BYTES + 256 + 1024; }
};

template <int DK, bool DROP>
__global__ void __launch_bounds__(ATT_THREADS, 1) attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                  const __grid_constant__ CUtensorMap tmK,
                                                                  const __grid_constant__ CUtensorMap tmV,
                                                                  const __grid_constant__ CUtensorMap tmO,
                                                                  const uint8_t* __restrict__ mask,
                                                                  float* __restrict__ stat_max,
                                                                  float* __restrict__ stat_sum, int S, int n_heads,
                                                                  float scale_log2e, DropSite drop,
                                                                  const int* __restrict__ extent,
                                                                  const int* __restrict__ pack_off) {
  (void)extent; (void)pack_off;   // the single-pass kernel always covers the whole (dense) slate
  using L = AttFwdSmem<DK>;
  constexpr int NKB = L::NKB;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + L::Q_BYTES;
  uint8_t* v_s = k_s + L::K_BYTES;
  uint8_t* o_s = v_s + L::V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_s + L::O_BYTES);
  uint64_t* load_bar = bars;        // Q, K, V landed
  uint64_t* s_bar = bars + 1;       // S = Q K^T complete
  uint64_t* p_bar = bars + 2;       // P written to TMEM by all 128 softmax threads
  uint64_t* o_bar = bars + 3;       // O = P V complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  uint32_t* mask_bits = tmem_slot + 2;   // 8 words: bit j of word w = key 32w+j is a real (unpadded, < S) item

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int S16 = (S + 15) & ~15;       // UMMA N of the score tile
  const int S8 = (S + 7) & ~7;          // contraction length of P V

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmO);
    ptx::mbar_init(load_bar, 1);
    ptx::mbar_init(s_bar, 1);
    ptx::mbar_init(p_bar, 128);
    ptx::mbar_init(o_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  arb_pdl_wait();
  if (warp >= 2) {
    const int et = threadIdx.x - 64;
    if (et < 8) {
      uint32_t w = 0;
      for (int j = 0; j < 32; ++j) {
        const int key = 32 * et + j;
        if (key < S && mask[size_t(b) * S + key] == 0) w |= (1u << j);
      }
      mask_bits[et] = w;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;            // columns [0, 256)
  const uint32_t tmem_O = tmem_base + 256;      // columns [256, 256 + DK)

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(load_bar, L::Q_BYTES + L::K_BYTES + L::V_BYTES);
      for (int kb = 0; kb < NKB; ++kb) {
        ptx::tma_load_4d(q_s + kb * (128 * 128), &tmQ, load_bar, 32 * kb, m0, head, b);
        ptx::tma_load_4d(k_s + kb * (256 * 128), &tmK, load_bar, 32 * kb, 0, head, b);
        ptx::tma_load_4d(v_s + kb * (256 * 128), &tmV, load_bar, 32 * kb, 0, head, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_wait(load_bar, 0);
      ptx::tc_fence_after();
      // ---- S = Q K^T
      const uint32_t idesc_s = ptx::idesc_tf32(128, S16, 0, 0);
      uint32_t acc = 0;
      for (int kb = 0; kb < NKB; ++kb) {
        const int ksteps = (DK - 32 * kb >= 32) ? 4 : (DK - 32 * kb + 7) / 8;
        const uint32_t qa = ptx::smem_u32(q_s + kb * (128 * 128));
        const uint32_t ka = ptx::smem_u32(k_s + kb * (256 * 128));
        for (int k = 0; k < ksteps; ++k) {
          ptx::mma_tf32_ss(tmem_S, ptx::smem_desc_sw128<2>(qa + k * 32, 16, 1024),
                           ptx::smem_desc_sw128<2>(ka + k * 32, 16, 1024), idesc_s, acc);
          acc = 1;
        }
      }
      ptx::mma_commit(s_bar);
      // ---- O = P V   (A = P from TMEM, B = V MN-major: 8 key rows = 1024 B per K step, 4-row swizzle atoms)
      ptx::mbar_wait(p_bar, 0);
      ptx::tc_fence_after();
      const uint32_t idesc_o = ptx::idesc_tf32(128, DK < 16 ? 16 : DK, 0, 1);
      const uint32_t va = ptx::smem_u32(v_s);
      for (int i = 0; i < S8 / 8; ++i) {
        ptx::mma_tf32_ts(tmem_O, tmem_S + 8 * i, ptx::smem_desc_sw128<1>(va + i * 1024, 256 * 128, 512), idesc_o,
                         i > 0 ? 1u : 0u);
      }
      ptx::mma_commit(o_bar);
    }
  } else {
    // ===================== softmax + epilogue: thread = query row =====================
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const int qidx = m0 + row;
    const uint32_t lane_addr = uint32_t(32 * q) << 16;
    ptx::mbar_wait(s_bar, 0);
    ptx::tc_fence_after();
    const int nchunks = (S8 + 31) / 32;
    float mx = -CUDART_INF_F;
    for (int c = 0; c < nchunks; ++c) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_S + lane_addr + 32 * c, v);
      ptx::tmem_ld_wait();
      const uint32_t bits = mask_bits[c];
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (bits & (1u << j)) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    float sum = 0.f;
    const float mxs = mx * scale_log2e;
    for (int c = 0; c < nchunks; ++c) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_S + lane_addr + 32 * c, v);
      ptx::tmem_ld_wait();
      const uint32_t bits = mask_bits[c];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        // exp((s - max)/sqrt(dk)) as exp2; padded keys contribute exactly 0.  An all-padded slate gives
        // (-inf) - (-inf) = NaN like the reference (quirk Q2).
        float e = (bits & (1u << j)) ? ex2_approx(fmaf(__uint_as_float(v[j]), scale_log2e, -mxs)) : 0.0f;
        sum += e;                       // softmax normalises BEFORE dropout (transformer.py:153-155)
        if constexpr (DROP) {
          const unsigned long long idx =
              ((unsigned long long)(b * n_heads + head) * S + qidx) * (unsigned long long)S + (32 * c + j);
          e = drop_keep(idx, drop.seed, drop.thresh) ? e * drop.scale : 0.0f;
        }
        v[j] = round_tf32(e);
      }
      ptx::tmem_st_32x32(tmem_S + lane_addr + 32 * c, v);
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    ptx::mbar_arrive(p_bar);
    if (qidx < S) {
      const size_t so = (size_t(b) * n_heads + head) * S + qidx;
      stat_max[so] = mx;
      stat_sum[so] = sum;
    }
    const float inv = 1.0f / sum;
    ptx::mbar_wait(o_bar, 0);
    ptx::tc_fence_after();
#pragma unroll
    for (int sl = 0; sl < NKB; ++sl) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_O + lane_addr + 32 * sl, v);
      ptx::tmem_ld_wait();
      uint8_t* slab_row = o_s + sl * (128 * 128) + row * 128;
#pragma unroll
      for (int piece = 0; piece < 8; ++piece) {
        float4 o;
        o.x = __uint_as_float(v[piece * 4 + 0]) * inv;
        o.y = __uint_as_float(v[piece * 4 + 1]) * inv;
        o.z = __uint_as_float(v[piece * 4 + 2]) * inv;
        o.w = __uint_as_float(v[piece * 4 + 3]) * inv;
        *reinterpret_cast<float4*>(slab_row + ((piece ^ (row & 7)) << 4)) = o;
      }
    }
    ptx::fence_proxy_async_smem();
    ptx::named_bar_sync(1, 128);
    if (threadIdx.x == 64) {
      for (int sl = 0; sl < NKB; ++sl) ptx::tma_store_4d(&tmO, o_s + sl * (128 * 128), 32 * sl, m0, head, b);
      ptx::tma_store_commit();
      ptx::tma_store_wait_all();
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------ two-pass variant
// Same result, organised so that TWO CTAs fit on an SM (the kernel above needs S(<=256)+O columns = a 512-column TMEM
// allocation, i.e. one CTA per SM, and is latency-bound at 7 % tensor pipe).  Here the score tile is produced in
// 128-key chunks into a 128-column TMEM window, twice: pass A only takes the row maximum, pass B recomputes the chunk
// (K = dk, a handful of MMAs), exponentiates against the final maximum, writes P in place and accumulates O += P V.
// TMEM: S chunk [0,128) + O [128,128+DK) -> 256 columns; shared memory 96 KB -> two co-resident CTAs overlap each
// other's load / MMA / softmax phases.  Head width <= 32.  Eight softmax warps per CTA: a warp may only touch the 32
// TMEM lanes of its quadrant, so two warps share each quadrant and split every 32-key chunk 16 / 16; the two partial
// row maxima (after pass A) and row sums (after pass B) are combined through shared memory.
template <int DK, bool DROP, bool OUT16 = false>
__global__ void __launch_bounds__(ATT2_THREADS, 2) attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV,
                                                                   const __grid_constant__ CUtensorMap tmO,
                                                                   const uint8_t* __restrict__ mask,
                                                                   float* __restrict__ stat_max,
                                                                   float* __restrict__ stat_sum, int S, int n_heads,
                                                                   float scale_log2e, DropSite drop,
                                                                   const int* __restrict__ extent,
                                                                   const int* __restrict__ pack_off) {
  static_assert(DK <= 32, "two-pass forward kernel: head width <= 32");
  using L = AttFwdSmem<DK>;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + L::Q_BYTES;
  uint8_t* v_s = k_s + L::K_BYTES;
  uint8_t* o_s = v_s + L::V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_s + L::O_BYTES);
  uint64_t* load_bar = bars;
  uint64_t* s_bar = bars + 1;       // a score chunk is complete            (one phase per step)
  uint64_t* t_bar = bars + 2;       // the 256 softmax threads are done with the chunk (one phase per step)
  uint64_t* o_bar = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  uint32_t* mask_bits = tmem_slot + 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  // Packed rows (pack_off != nullptr): the activations hold only the first round_up(extent, 16) rows of every slate,
  // slate b starting at row pack_off[b] of one long [rows, h, dk] tensor (TMA coordinates (.., row, head, 0)).  Boxes
  // that overrun the slate read the next slates' rows (finite; their keys are masked, their query rows never stored);
  // the output is stored in 16-row boxes that stop at the slate's last packed row.  extent / pack_off were written by
  // kernels at least two launches upstream, so they may be read before the PDL wait.
  const bool packed = pack_off != nullptr;
  const int row_base = packed ? __ldg(pack_off + b) : 0;
  const int bc = packed ? 0 : b;
  if (packed && m0 >= ((__ldg(extent + b) + 15) & ~15)) return;   // no packed query rows in this tile (or an empty slate)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmO);
    ptx::mbar_init(load_bar, 1);
    ptx::mbar_init(s_bar, 1);
    ptx::mbar_init(t_bar, ATT2_SOFTMAX);
    ptx::mbar_init(o_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<256>(tmem_slot);
  arb_pdl_wait();
  // keys at or beyond the slate's extent are all masked (probability exactly 0): the score / P V products, the softmax
  // loop and the K / V loads stop there.  At least one key column group is always processed, so an all-padded slate
  // still produces the reference's NaN rows.
  const int kext = extent ? max(1, min(S, __ldg(extent + b))) : S;
  const int S16 = (kext + 15) & ~15, S8 = (kext + 7) & ~7;
  const int nkc = (S16 + 127) / 128;          // key chunks
  const int nsteps = 1 + nkc;                 // one pass-A step over the whole score row + one pass-B step per chunk
  if (warp == 0 && lane == 0) {
    // The loads go out right here -- the thread that initialised the barriers needs nobody else -- so that they run
    // behind the TMEM allocation, the mask read and the block-wide barrier below instead of after them.
    // K and V arrive as 128-key boxes: only the chunks that hold keys below the extent are fetched
    ptx::mbar_expect_tx(load_bar, L::Q_BYTES + nkc * 2 * (128 * 128));
    ptx::tma_load_4d(q_s, &tmQ, load_bar, 0, row_base + m0, head, bc);
    for (int kc = 0; kc < nkc; ++kc) {
      ptx::tma_load_4d(k_s + kc * 16384, &tmK, load_bar, 0, row_base + 128 * kc, head, bc);
      ptx::tma_load_4d(v_s + kc * 16384, &tmV, load_bar, 0, row_base + 128 * kc, head, bc);
    }
  }
  if (warp >= 2) {      // key mask as 8 words: one key per softmax thread, one ballot per warp
    const int key = threadIdx.x - 64;
    const uint32_t w = __ballot_sync(0xffffffffu, key < S && mask[size_t(b) * S + key] == 0);
    if (lane == 0) mask_bits[warp - 2] = w;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    // (the producer's loads were issued above)
  } else if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_wait(load_bar, 0);
      ptx::tc_fence_after();
      constexpr int KSTEPS = (DK + 7) / 8;
      const uint32_t qa = ptx::smem_u32(q_s), ka = ptx::smem_u32(k_s), va = ptx::smem_u32(v_s);
      const uint32_t idesc_o = ptx::idesc_tf32(128, DK < 16 ? 16 : DK, 0, 1);
      auto issue_pv = [&](int kc) {
        const int keys = min(128, S8 - 128 * kc);
        for (int i = 0; i < keys / 8; ++i)
          ptx::mma_tf32_ts(tmem_O, tmem_S + 8 * i, ptx::smem_desc_sw128<1>(va + kc * 16384 + i * 1024, 256 * 128, 512),
                           idesc_o, (kc > 0 || i > 0) ? 1u : 0u);
      };
      // step 0 (pass A): the WHOLE score row block S = Q K^T, N = S16 <= 256 columns, in one issue -- it may spill over
      // the O columns, which pass B only starts to use after every softmax thread has taken its row maximum
      {
        const uint32_t idesc_a = ptx::idesc_tf32(128, S16, 0, 0);
        for (int k = 0; k < KSTEPS; ++k)
          ptx::mma_tf32_ss(tmem_S, ptx::smem_desc_sw128<2>(qa + k * 32, 16, 1024),
                           ptx::smem_desc_sw128<2>(ka + k * 32, 16, 1024), idesc_a, k > 0);
        ptx::mma_commit(s_bar);
      }
      // steps 1..nkc (pass B): recompute one 128-key chunk into the 128-column window; the previous chunk's P V runs
      // first (P sits in the window until then)
      for (int st = 1; st < nsteps; ++st) {
        const int kc = st - 1;
        ptx::mbar_wait(t_bar, (st - 1) & 1);
        ptx::tc_fence_after();
        if (kc > 0) issue_pv(kc - 1);
        const int nc = min(128, S16 - 128 * kc);
        const uint32_t idesc_s = ptx::idesc_tf32(128, nc, 0, 0);
        for (int k = 0; k < KSTEPS; ++k)
          ptx::mma_tf32_ss(tmem_S, ptx::smem_desc_sw128<2>(qa + k * 32, 16, 1024),
                           ptx::smem_desc_sw128<2>(ka + kc * 16384 + k * 32, 16, 1024), idesc_s, k > 0);
        ptx::mma_commit(s_bar);
      }
      ptx::mbar_wait(t_bar, (nsteps - 1) & 1);
      ptx::tc_fence_after();
      issue_pv(nkc - 1);
      ptx::mma_commit(o_bar);
    }
  } else {
    const int q = warp & 3;                     // TMEM lane quadrant
    const int sub = (warp - 2) >> 2;            // which 16 keys of every 32-key chunk this warp handles
    const int row = 32 * q + lane;
    const int qidx = m0 + row;
    const uint32_t lane_addr = uint32_t(32 * q) << 16;
    float* xch_max = reinterpret_cast<float*>(o_s);     // [2][128] partial maxima (the O staging is idle until the end)
    float* xch_sum = reinterpret_cast<float*>(k_s);     // [2][128] partial sums   (K is dead after the last score chunk)
    float mx = -CUDART_INF_F, sum = 0.f, mxs = 0.f;
    for (int st = 0; st < nsteps; ++st) {
      const bool pass_b = st > 0;
      const int kc = pass_b ? st - 1 : 0;
      if (st == 1) {                            // pass A done: combine the two partial row maxima
        xch_max[sub * 128 + row] = mx;
        ptx::named_bar_sync(1, ATT2_SOFTMAX);
        mx = fmaxf(xch_max[row], xch_max[128 + row]);
        mxs = mx * scale_log2e;
      }
      ptx::mbar_wait(s_bar, st & 1);
      ptx::tc_fence_after();
      // pass A walks every 32-key group of the row block, pass B the groups of its chunk
      const int keys = pass_b ? min(128, S8 - 128 * kc) : S8;
      const int nch = (keys + 31) / 32;
      for (int c = 0; c < nch; ++c) {
        uint32_t v[16];
        const int col = 32 * c + 16 * sub;
        ptx::tmem_ld_32x16(tmem_S + lane_addr + col, v);
        ptx::tmem_ld_wait();
        const uint32_t bits = mask_bits[4 * kc + c] >> (16 * sub);
        if (!pass_b) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (bits & (1u << j)) mx = fmaxf(mx, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float e = (bits & (1u << j)) ? ex2_approx(fmaf(__uint_as_float(v[j]), scale_log2e, -mxs)) : 0.0f;
            sum += e;
            if constexpr (DROP) {
              const unsigned long long idx =
                  ((unsigned long long)(b * n_heads + head) * S + qidx) * (unsigned long long)S + (128 * kc + col + j);
              e = drop_keep(idx, drop.seed, drop.thresh) ? e * drop.scale : 0.0f;
            }
            v[j] = round_tf32(e);
          }
          ptx::tmem_st_32x16(tmem_S + lane_addr + col, v);
        }
      }
      if (pass_b) ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(t_bar);
    }
    xch_sum[sub * 128 + row] = sum;             // every score MMA has completed: the K tile is free
    ptx::named_bar_sync(1, ATT2_SOFTMAX);
    sum = xch_sum[row] + xch_sum[128 + row];
    if (sub == 0 && qidx < S) {
      const size_t so = (size_t(b) * n_heads + head) * S + qidx;
      stat_max[so] = mx;
      stat_sum[so] = sum;
    }
    const float inv = 1.0f / sum;
    ptx::mbar_wait(o_bar, 0);
    ptx::tc_fence_after();
    {
      uint32_t v[16];                           // this warp's 16 of the (up to) 32 output columns
      ptx::tmem_ld_32x16(tmem_O + lane_addr + 16 * sub, v);
      ptx::tmem_ld_wait();
      if constexpr (OUT16) {
        // bf16 mode: the context only feeds the output projection -- stage it as dense bfloat16 rows (32 columns =
        // 64 bytes, unswizzled tensor map); this warp's 16 columns are bytes [32 sub, 32 sub + 32)
        uint4* dst = reinterpret_cast<uint4*>(o_s + row * 64 + sub * 32);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          uint4 pk;
          pk.x = ptx::pack_bf16(__uint_as_float(v[8 * k + 0]) * inv, __uint_as_float(v[8 * k + 1]) * inv);
          pk.y = ptx::pack_bf16(__uint_as_float(v[8 * k + 2]) * inv, __uint_as_float(v[8 * k + 3]) * inv);
          pk.z = ptx::pack_bf16(__uint_as_float(v[8 * k + 4]) * inv, __uint_as_float(v[8 * k + 5]) * inv);
          pk.w = ptx::pack_bf16(__uint_as_float(v[8 * k + 6]) * inv, __uint_as_float(v[8 * k + 7]) * inv);
          dst[k] = pk;
        }
      } else {
        uint8_t* slab_row = o_s + row * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int piece = 4 * sub + k;
          float4 o;
          o.x = __uint_as_float(v[k * 4 + 0]) * inv;
          o.y = __uint_as_float(v[k * 4 + 1]) * inv;
          o.z = __uint_as_float(v[k * 4 + 2]) * inv;
          o.w = __uint_as_float(v[k * 4 + 3]) * inv;
          *reinterpret_cast<float4*>(slab_row + ((piece ^ (row & 7)) << 4)) = o;
        }
      }
    }
    ptx::fence_proxy_async_smem();
    ptx::named_bar_sync(1, ATT2_SOFTMAX);
    if (threadIdx.x == 64) {
      if (packed) {       // 16-row boxes up to the slate's last packed row (the staged rows are 128 / 64 bytes wide)
        const int n16 = (min(128, S16 - m0) + 15) >> 4;
        for (int i = 0; i < n16; ++i)
          ptx::tma_store_4d(&tmO, o_s + i * (OUT16 ? 1024 : 2048), 0, row_base + m0 + 16 * i, head, 0);
      } else {
        ptx::tma_store_4d(&tmO, o_s, 0, m0, head, b);
      }
      ptx::tma_store_commit();
      ptx::tma_store_wait_read();
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<256>(tmem_base);
  }
}

static int g_attn_fwd_two_pass = 1;
void set_attn_fwd_two_pass(int on) { g_attn_fwd_two_pass = on; }

template <int DK>
static int launch_fwd_t(const AttnFwdArgs& a, cudaStream_t st) {
  using L = AttFwdSmem<DK>;
  alignas(64) CUtensorMap tQ, tK, tV, tO;
  int rc;
  if ((rc = make_tmap_4d(&tQ, a.q, TmapBox{{32, 128, 1, 1}}, 0, 1))) return rc;
  const uint32_t kv_rows = (DK <= 32 && g_attn_fwd_two_pass) ? 128u : 256u;   // two-pass kernel: one box per key chunk
  if ((rc = make_tmap_4d(&tK, a.k, TmapBox{{32, kv_rows, 1, 1}}, 0, 1))) return rc;
  if ((rc = make_tmap_4d(&tV, a.v, TmapBox{{32, kv_rows, 1, 1}}, 1, 1))) return rc;
  const bool out16 = a.o.bf16 != 0;
  if (out16 && !(DK <= 32 && g_attn_fwd_two_pass)) { arb_set_error("attn_fwd: a bf16 context needs the two-pass kernel (head width <= 32)"); return ARB_E_UNSUPPORTED; }
  const bool packed = a.pack_off != nullptr;
  if (packed && !(DK <= 32 && g_attn_fwd_two_pass && a.extent)) { arb_set_error("attn_fwd: packed rows need the two-pass kernel and the slate extents"); return ARB_E_UNSUPPORTED; }
  if ((rc = make_tmap_4d(&tO, a.o, TmapBox{{32, packed ? 16u : 128u, 1, 1}}, out16 ? 2 : 0, 0))) return rc;
  const bool drop = a.drop.thresh != 0;
  void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, const uint8_t*, float*, float*, int, int, float,
               DropSite, const int*, const int*);
  if constexpr (DK <= 32) {
    if (g_attn_fwd_two_pass) {
      if (out16) kern = drop ? attn_fwd2_kernel<DK, true, true> : attn_fwd2_kernel<DK, false, true>;
      else kern = drop ? attn_fwd2_kernel<DK, true> : attn_fwd2_kernel<DK, false>;
    } else {
      kern = drop ? attn_fwd_kernel<DK, true> : attn_fwd_kernel<DK, false>;
    }
  } else {
    kern = drop ? attn_fwd_kernel<DK, true> : attn_fwd_kernel<DK, false>;
  }
  static bool configured[ARB_MAX_DEVICES][8] = {};
  const int slot = (drop ? 1 : 0) + ((DK <= 32 && g_attn_fwd_two_pass) ? 2 : 0) + (out16 ? 4 : 0);
  const int dev = arb_device_slot();
  if (!configured[dev][slot]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total()) != cudaSuccess) {
      arb_set_error("attn_fwd: cannot raise the dynamic shared memory limit");
      return ARB_E_CUDA;
    }
    configured[dev][slot] = true;
  }
  dim3 grid((a.S + 127) / 128, a.h, a.B);
  {
    ProfScope ps(ARB_PROF_GEMM, (a.extent ? arb_attn_frac() : 1.0) * 4.0 * double(a.S) * a.S * a.dk * a.h * a.B, st,
                 (packed ? arb_row_frac() : 1.0) * 4.0 * double(a.B) * a.h * a.S * ((out16 ? 3.5 : 4.0) * a.dk + 2.0),
                 (DK <= 32 && g_attn_fwd_two_pass) ? "attn_fwd2_kernel" : "attn_fwd_kernel");
    const int threads = (DK <= 32 && g_attn_fwd_two_pass) ? ATT2_THREADS : ATT_THREADS;
    arb_launch(kern, grid, dim3(threads), size_t(L::total()), st, tQ, tK, tV, tO, a.mask, a.stat_max, a.stat_sum, a.S, a.h,
               a.scale * 1.4426950408889634f, a.drop, a.extent, a.pack_off);
  }
  arb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { arb_set_error(cudaGetErrorString(e)); return ARB_E_CUDA; }
  return ARB_OK;
}

bool attn_fused_supported(int S, int dk) { return S >= 1 && S <= 256 && (dk == 16 || dk == 32 || dk == 64); }

int launch_attn_fwd(const AttnFwdArgs& a, cudaStream_t st) {
  if (!attn_fused_supported(a.S, a.dk)) { arb_set_error("fused attention: unsupported shape"); return ARB_E_UNSUPPORTED; }
  switch (a.dk) {
    case 16: return launch_fwd_t<16>(a, st);
    case 32: return launch_fwd_t<32>(a, st);
    default: return launch_fwd_t<64>(a, st);
  }
}

}  // namespace arb
