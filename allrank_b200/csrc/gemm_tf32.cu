// Batched TF32 GEMM on the 5th-generation tensor cores:  C[M,N] = epilogue(alpha * A * B^T)
//
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a STAGES-deep shared-memory ring,
//   * tcgen05.mma.kind::tf32 issued by one elected thread, fp32 accumulator tile [128 x BLOCK_N] in TMEM,
//   * epilogue warps read TMEM (tcgen05.ld 32x32b), apply bias / ReLU / residual / ReLU-mask, stage the tile in
//     swizzled shared memory and write it with one TMA store per 32-column slab (TMA clips ragged edges, so a
//     240-row slate or a 136-wide feature matrix needs no masking code),
//   * either operand may be "K-major" (rows of K contiguous, e.g. nn.Linear weights [out,in]) or "MN-major"
//     (the transpose), which is what the backward GEMMs dX = dY W and dW = dY^T X need,
//   * split-K with red.global.add for the weight gradients (K = all rows of the batch).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue
// (warp w owns TMEM lanes 32*(w%4) .. +31, i.e. 32 rows of the tile).
//
// This one kernel is the tensor-core workhorse of the scorer: every nn.Linear forward/backward and the
// generic (unfused) attention contractions go through it.  Reference ops replaced: aten::addmm / aten::bmm
// issued from allrank/models/transformer.py:148,156,193-195,203,227 and allrank/models/model.py:41-44,117.
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.h"
#include "defaults.h"
#include "gemm_tf32.h"
#include "sm100_ptx.cuh"

namespace arb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;                    // 32 fp32 = one 128-byte swizzle row
constexpr int UMMA_K = 8;                      // kind::tf32: 8 elements (32 bytes) per instruction
// bf16 operands (kind::f16) keep the BYTE geometry: a k-block is still one 128-byte swizzle row (64 bf16), an
// instruction still consumes 32 bytes of K (16 bf16), so ring stages, K-major descriptors and k-step advances are
// shared; only MN-major tiles differ (plain SWIZZLE_128B slabs of 64 MN-elements x 64 k-rows instead of tf32's
// 32-byte-atom slabs of 32 x 32).
constexpr int BLOCK_K_BF16 = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 4;   // 16 KB
constexpr int GEMM_EPI_GROUPS = 1;            // epilogue warps per TMEM lane quadrant (2 measured slower: fewer CTAs per SM)
constexpr int GEMM_EPI_THREADS = 128 * GEMM_EPI_GROUPS;
constexpr int GEMM_THREADS = 64 + GEMM_EPI_THREADS;

struct GemmParams {
  int M, N, K;
  int nb2;
  int a_b2, a_b3, b_b2, b_b3, c_b2, c_b3;
  int flags;
  int kb_per_split;   // k-blocks per split (split-K), or total k-blocks
  float alpha;
  const float* bias;
  float* atomic_out;
  long long atomic_ld;
  DropSite drop;
  float* colsum_out;
  const int* rows_dev;   // packed rows: device-resident live row count -- bounds M, or K for the split-K weight gradients
  uint32_t* bits;        // EPI_RELU_BITS / EPI_MASK_BITS: [M, N / 32] words
};

// ---- epilogue element transform -----------------------------------------------------------------------------------
// One thread turns its 32 accumulator values of a 32-column chunk into the staged fp32 output row piece by piece.
// The flags are COMPILE-TIME here: with run-time flags the unrolled loop carries a test per element and flag (about
// 20 instructions per element), and the epilogue -- not HBM -- bounds the short-K linears (ncu, round 2: the eight
// epilogue warps of the persistent kernel issue 37 % of all cycles at 31 % DRAM utilisation).  epi_chunk_dispatch
// picks the instantiation once per chunk; unusual flag combinations take the generic run-time path.
// `word`: the chunk's 32 mask bits of this row (EPI_MASK_BITS); returns the chunk's ReLU bits (EPI_RELU_BITS, else 0).
template <int F>
__device__ __forceinline__ uint32_t epi_chunk_f32(const uint32_t (&v)[32], uint8_t* slab_row, int row, const float* bias_c,
                                                  float alpha, uint32_t word = 0u) {
  uint32_t out_bits = 0u;
#pragma unroll
  for (int piece = 0; piece < 8; ++piece) {
    float4* dst = reinterpret_cast<float4*>(slab_row + ((piece ^ (row & 7)) << 4));
    float4 o = make_float4(__uint_as_float(v[piece * 4 + 0]) * alpha, __uint_as_float(v[piece * 4 + 1]) * alpha,
                           __uint_as_float(v[piece * 4 + 2]) * alpha, __uint_as_float(v[piece * 4 + 3]) * alpha);
    if constexpr ((F & EPI_BIAS) != 0) {
      const float4 bv = *reinterpret_cast<const float4*>(bias_c + 4 * piece);
      o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
    }
    if constexpr ((F & EPI_RELU) != 0) {
      o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f);
    }
    if constexpr ((F & (EPI_ADD_AUX | EPI_MASK_AUX)) != 0) {
      const float4 a = *dst;
      if constexpr ((F & EPI_ADD_AUX) != 0) { o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
      if constexpr ((F & EPI_MASK_AUX) != 0) {
        o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f; o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f;
      }
    }
    if constexpr ((F & EPI_MASK_BITS) != 0) {
      o.x = (word >> (4 * piece + 0)) & 1u ? o.x : 0.f; o.y = (word >> (4 * piece + 1)) & 1u ? o.y : 0.f;
      o.z = (word >> (4 * piece + 2)) & 1u ? o.z : 0.f; o.w = (word >> (4 * piece + 3)) & 1u ? o.w : 0.f;
    }
    if constexpr ((F & EPI_RELU_BITS) != 0) {
      out_bits |= (o.x > 0.f ? 1u : 0u) << (4 * piece + 0) | (o.y > 0.f ? 1u : 0u) << (4 * piece + 1) |
                  (o.z > 0.f ? 1u : 0u) << (4 * piece + 2) | (o.w > 0.f ? 1u : 0u) << (4 * piece + 3);
    }
    *dst = o;
  }
  return out_bits;
}
// returns false when the flag combination has no specialisation (caller runs the generic loop); `bits_at`: this row's
// word of the chunk in the bit-mask array (written for EPI_RELU_BITS; nullptr: out of range); word_in: the same word,
// already loaded, for EPI_MASK_BITS
__device__ __forceinline__ bool epi_chunk_dispatch(int flags, const uint32_t (&v)[32], uint8_t* slab_row, int row,
                                                   const float* bias_c, float alpha, uint32_t* bits_at = nullptr,
                                                   uint32_t word_in = 0u) {
  switch (flags & (EPI_BIAS | EPI_RELU | EPI_ADD_AUX | EPI_MASK_AUX | EPI_DROPOUT | EPI_ATOMIC | EPI_RELU_BITS | EPI_MASK_BITS)) {
    case 0: epi_chunk_f32<0>(v, slab_row, row, bias_c, alpha); return true;
    case EPI_BIAS: epi_chunk_f32<EPI_BIAS>(v, slab_row, row, bias_c, alpha); return true;
    case EPI_BIAS | EPI_RELU: epi_chunk_f32<EPI_BIAS | EPI_RELU>(v, slab_row, row, bias_c, alpha); return true;
    case EPI_BIAS | EPI_ADD_AUX: epi_chunk_f32<EPI_BIAS | EPI_ADD_AUX>(v, slab_row, row, bias_c, alpha); return true;
    case EPI_MASK_AUX: epi_chunk_f32<EPI_MASK_AUX>(v, slab_row, row, bias_c, alpha); return true;
    case EPI_BIAS | EPI_RELU | EPI_RELU_BITS: {
      const uint32_t w = epi_chunk_f32<EPI_BIAS | EPI_RELU | EPI_RELU_BITS>(v, slab_row, row, bias_c, alpha);
      if (bits_at) *bits_at = w;
      return true;
    }
    case EPI_MASK_BITS:
      epi_chunk_f32<EPI_MASK_BITS>(v, slab_row, row, bias_c, alpha, word_in);   // (fetched one chunk ahead by the caller)
      return true;
    default: return false;
  }
}

template <int BLOCK_N, int NST = 2>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGING_BYTES = BLOCK_M * BLOCK_N * 4;
  // Two stages: the contraction is short (K = 32..512) and latency is hidden by co-resident CTAs instead
  // (~80 KB per CTA with an epilogue tile -> 2 per SM; ~50 KB without -> 4 per SM).
  // weight-gradient GEMMs (both operands MN-major, K = all rows of the batch) run a long K loop per CTA: 4 stages
  static constexpr int stages() { return NST; }
  static constexpr int pipe_bytes() { return stages() * STAGE_BYTES; }
  // without an aux (residual / mask) tile the output staging aliases the operand ring: it is only written after
  // the last MMA has consumed the ring
  // The output staging tile (and the residual / mask tile, fetched only after the last MMA) aliases the operand
  // ring: both are touched only once every MMA of the CTA has completed.
  static constexpr int body_bytes() { return pipe_bytes() > STAGING_BYTES ? pipe_bytes() : STAGING_BYTES; }
  static constexpr int total() { return body_bytes() + 256 + BLOCK_N * 4 + 1024; }
};

template <int BLOCK_N, int A_MN, int B_MN, bool DROP, int NST, bool IN16 = false, bool OUT16 = false>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const __grid_constant__ CUtensorMap tmC,
                                                                 const __grid_constant__ CUtensorMap tmAux,
                                                                 const GemmParams p) {
  using L = SmemLayout<BLOCK_N, NST>;
  constexpr int STAGES = L::stages();
  constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr int N_SLABS = BLOCK_N / 32;                       // 32-column accumulator chunks
  constexpr int OUT_COLS = OUT16 ? 64 : 32;                   // output columns per 128-byte staging slab row
  constexpr int OUT_SLABS = (BLOCK_N + OUT_COLS - 1) / OUT_COLS;
  static_assert(!OUT16 || BLOCK_N >= 64, "bf16 outputs need at least one full 128-byte slab row");

  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const bool has_aux = (p.flags & (EPI_ADD_AUX | EPI_MASK_AUX)) != 0;
  uint8_t* staging = smem;
  uint8_t* tail = smem + L::body_bytes();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* aux_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 1);
  float* bias_s = reinterpret_cast<float*>(tail + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_N, m0 = blockIdx.y * BLOCK_M;
  const bool split = (p.flags & EPI_ATOMIC) != 0;
  const int b2 = split ? 0 : int(blockIdx.z) % p.nb2, b3 = split ? 0 : int(blockIdx.z) / p.nb2;
  constexpr int KELEMS = IN16 ? BLOCK_K_BF16 : BLOCK_K;       // elements of K per k-block (128 bytes either way)
  constexpr int MN_SLAB = IN16 ? 64 : 32;                     // MN-elements per MN-major slab (128 bytes)
  constexpr int MN_SLAB_BYTES = MN_SLAB * KELEMS * (IN16 ? 2 : 4);   // 8192 (bf16) / 4096 (tf32)
  // Packed rows: the live row count is on the device (written at least two launches upstream, so it may be read
  // before the PDL wait).  It bounds M -- CTAs of tiles beyond it leave at once -- or, for the split-K weight gradients,
  // K, which is then divided evenly over the grid's splits here.
  const bool dev_rows = p.rows_dev != nullptr;
  if (dev_rows && !split && m0 >= __ldg(p.rows_dev)) return;
  const int k_live = (dev_rows && split) ? min(p.K, __ldg(p.rows_dev)) : p.K;
  const int total_kb = (k_live + KELEMS - 1) / KELEMS;
  const int kb_per = (dev_rows && split) ? (total_kb + int(gridDim.z) - 1) / int(gridDim.z) : p.kb_per_split;
  const int kb_begin = split ? int(blockIdx.z) * kb_per : 0;
  const int kb_end = split ? min(total_kb, kb_begin + kb_per) : total_kb;
  const int nkb = max(0, kb_end - kb_begin);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (!split) ptx::prefetch_tmap(&tmC);
    if (has_aux) ptx::prefetch_tmap(&tmAux);
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::mbar_init(aux_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  arb_pdl_wait();          // everything above overlaps the previous kernel's tail; global memory is touched below
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, round = i / STAGES;
        if (round > 0) ptx::mbar_wait(&empty_bar[s], (round - 1) & 1);
        uint8_t* a_s = smem + s * L::STAGE_BYTES;
        uint8_t* b_s = a_s + A_STAGE_BYTES;
        const int k0 = (kb_begin + i) * KELEMS;
        ptx::mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        if (A_MN) {
          for (int c = 0; c < BLOCK_M / MN_SLAB; ++c)
            ptx::tma_load_4d(a_s + c * MN_SLAB_BYTES, &tmA, &full_bar[s], m0 + MN_SLAB * c, k0, b2 * p.a_b2, b3 * p.a_b3);
        } else {
          ptx::tma_load_4d(a_s, &tmA, &full_bar[s], k0, m0, b2 * p.a_b2, b3 * p.a_b3);
        }
        if (B_MN) {
          for (int c = 0; c < BLOCK_N / MN_SLAB; ++c)
            ptx::tma_load_4d(b_s + c * MN_SLAB_BYTES, &tmB, &full_bar[s], n0 + MN_SLAB * c, k0, b2 * p.b_b2, b3 * p.b_b3);
        } else {
          ptx::tma_load_4d(b_s, &tmB, &full_bar[s], k0, n0, b2 * p.b_b2, b3 * p.b_b3);
        }
      }
      if (has_aux) {
        // the residual / ReLU-mask tile goes into the staging area, i.e. over the operand ring: wait until the
        // tensor core has finished reading it.  It has the output's element type: 128-byte slab rows hold 32 fp32 or
        // 64 bf16 columns.
        if (nkb > 0) ptx::mbar_wait(tmem_full_bar, 0);
        ptx::mbar_expect_tx(aux_bar, OUT16 ? L::STAGING_BYTES / 2 : L::STAGING_BYTES);
        for (int c = 0; c < OUT_SLABS; ++c)
          ptx::tma_load_4d(staging + c * (BLOCK_M * 128), &tmAux, aux_bar, n0 + OUT_COLS * c, m0, b2 * p.c_b2, b3 * p.c_b3);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = IN16 ? ptx::idesc_bf16(BLOCK_M, BLOCK_N, A_MN, B_MN)
                                      : ptx::idesc_tf32(BLOCK_M, BLOCK_N, A_MN, B_MN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, round = i / STAGES;
        ptx::mbar_wait(&full_bar[s], round & 1);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {      // 4 instructions per k-block for both element types
          // K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); advance 32 B per K step inside the swizzle row
          // MN-major tf32: 32-wide slabs of [32 k-rows x 128 B] 4096 B apart (LBO); swizzle atom = 4 k-rows (512 B,
          //           SBO); 8 k-rows (1024 B) per K step
          // MN-major bf16: 64-wide slabs of [64 k-rows x 128 B] 8192 B apart (LBO); 8-k-row groups 1024 B apart (SBO);
          //           16 k-rows (2048 B) per K step; plain SWIZZLE_128B
          uint64_t da, db;
          if constexpr (IN16) {
            da = A_MN ? ptx::smem_desc_sw128<2>(a_addr + k * 2048, 8192, 1024) : ptx::smem_desc_sw128<2>(a_addr + k * 32, 16, 1024);
            db = B_MN ? ptx::smem_desc_sw128<2>(b_addr + k * 2048, 8192, 1024) : ptx::smem_desc_sw128<2>(b_addr + k * 32, 16, 1024);
            ptx::mma_bf16_ss(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
          } else {
            da = A_MN ? ptx::smem_desc_sw128<1>(a_addr + k * 1024, 4096, 512) : ptx::smem_desc_sw128<2>(a_addr + k * 32, 16, 1024);
            db = B_MN ? ptx::smem_desc_sw128<1>(b_addr + k * 1024, 4096, 512) : ptx::smem_desc_sw128<2>(b_addr + k * 32, 16, 1024);
            ptx::mma_tf32_ss(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
          }
        }
        ptx::mma_commit(&empty_bar[s]);     // frees the stage when these MMAs have read it
      }
      ptx::mma_commit(tmem_full_bar);       // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..) =====================
    // One warp per TMEM lane quadrant walks the 32-column chunks of the accumulator.  (GEMM_EPI_GROUPS = 2 -- two
    // warps per quadrant on alternate chunks -- shortens a CTA's epilogue but costs a resident CTA per SM: the
    // short-K products with an aux tile went from 0.73 / 0.90 to 0.59 / 0.77 of the HBM roof; profiles/r2/call18.)
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = 32 * q + lane;          // row of the tile owned by this thread
    const int et = threadIdx.x - 64;        // 0..GEMM_EPI_THREADS-1
    const int grp = (warp - 2) >> 2;        // which chunks: grp, grp + GEMM_EPI_GROUPS, ...
    if (p.flags & EPI_BIAS) {
      for (int j = et; j < BLOCK_N; j += GEMM_EPI_THREADS) bias_s[j] = (n0 + j < p.N) ? p.bias[n0 + j] : 0.0f;
    }
    ptx::named_bar_sync(1, GEMM_EPI_THREADS);
    // EPI_MASK_BITS: this row's mask word of a chunk is fetched one chunk ahead (a global load per row and chunk); the
    // first one goes out before the wait for the accumulator
    auto mask_word = [&](int c) -> uint32_t {
      return ((p.flags & EPI_MASK_BITS) && c < N_SLABS && m0 + row < p.M && n0 + 32 * c < p.N)
                 ? p.bits[(long long)(m0 + row) * (p.N >> 5) + ((n0 >> 5) + c)] : 0u;
    };
    uint32_t word_next = mask_word(grp);
    if (nkb > 0) {
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
    }
    if (has_aux) ptx::mbar_wait(aux_bar, 0);
#pragma unroll 1
    for (int c = grp; c < N_SLABS; c += GEMM_EPI_GROUPS) {
      const uint32_t word_in = word_next;
      word_next = mask_word(c + GEMM_EPI_GROUPS);
      uint32_t v[32];
      if (nkb > 0) {
        ptx::tmem_ld_32x32(tmem_base + (uint32_t(32 * q) << 16) + uint32_t(32 * c), v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if constexpr (OUT16) {
        // bf16 output: a 128-byte staging row holds 64 columns; this 32-column chunk fills 16-byte pieces
        // 4*(c&1) .. +3 of slab c/2.  The aux (ReLU-mask) tile has the same layout and type.
        uint8_t* slab_row = staging + (c >> 1) * (BLOCK_M * 128) + row * 128;
#pragma unroll
        for (int piece = 0; piece < 4; ++piece) {
          uint4* dst = reinterpret_cast<uint4*>(slab_row + ((((c & 1) * 4 + piece) ^ (row & 7)) << 4));
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = piece * 8 + e;
            float x = __uint_as_float(v[j]) * p.alpha;
            if (p.flags & EPI_BIAS) x += bias_s[32 * c + j];
            if (p.flags & EPI_RELU) x = fmaxf(x, 0.0f);
            if constexpr (DROP) {
              const unsigned long long idx = (unsigned long long)(m0 + row) * (unsigned long long)p.N + (n0 + 32 * c + j);
              x = drop_keep(idx, p.drop.seed, p.drop.thresh) ? x * p.drop.scale : 0.0f;
            }
            o[e] = x;
          }
          if (has_aux) {                       // bf16 aux: bit 15 clear and non-zero <=> value > 0
            const uint4 a = *dst;
            const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t h = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
              const float av = __uint_as_float(h << 16);
              if (p.flags & EPI_ADD_AUX) o[e] += av;
              if (p.flags & EPI_MASK_AUX) o[e] = av > 0.f ? o[e] : 0.f;
            }
          }
          uint4 pk;
          pk.x = ptx::pack_bf16(o[0], o[1]); pk.y = ptx::pack_bf16(o[2], o[3]);
          pk.z = ptx::pack_bf16(o[4], o[5]); pk.w = ptx::pack_bf16(o[6], o[7]);
          *dst = pk;
        }
        continue;
      }
      uint8_t* slab_row = staging + c * (BLOCK_M * 128) + row * 128;
      if constexpr (!DROP) {
        uint32_t* bits_at = (p.bits && m0 + row < p.M && n0 + 32 * c < p.N)
                                ? p.bits + (long long)(m0 + row) * (p.N >> 5) + ((n0 >> 5) + c) : nullptr;
        if (epi_chunk_dispatch(p.flags, v, slab_row, row, bias_s + 32 * c, p.alpha, bits_at, word_in)) continue;
      }
#pragma unroll
      for (int piece = 0; piece < 8; ++piece) {
        float4* dst = reinterpret_cast<float4*>(slab_row + ((piece ^ (row & 7)) << 4));
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = piece * 4 + e;
          float x = __uint_as_float(v[j]) * p.alpha;
          if (p.flags & EPI_BIAS) x += bias_s[32 * c + j];
          if (p.flags & EPI_RELU) x = fmaxf(x, 0.0f);
          if constexpr (DROP) {
            const unsigned long long idx = (unsigned long long)(m0 + row) * (unsigned long long)p.N + (n0 + 32 * c + j);
            x = drop_keep(idx, p.drop.seed, p.drop.thresh) ? x * p.drop.scale : 0.0f;
          }
          o[e] = x;
        }
        if (has_aux) {
          const float4 a = *dst;
          if (p.flags & EPI_ADD_AUX) { o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w; }
          if (p.flags & EPI_MASK_AUX) {
            o[0] = a.x > 0.f ? o[0] : 0.f; o[1] = a.y > 0.f ? o[1] : 0.f;
            o[2] = a.z > 0.f ? o[2] : 0.f; o[3] = a.w > 0.f ? o[3] : 0.f;
          }
        }
        if (split) {
          // one 128-bit reduction per 4 columns: the weight-gradient GEMMs are bound by the number of L2 atomic
          // operations (measured: 4.8 M scalar atomics = ~120 us regardless of the bytes streamed)
          const int gm = m0 + row, gn = n0 + 32 * c + piece * 4;
          if (gm < p.M && gn < p.N) {
            float* dstg = p.atomic_out + (long long)gm * p.atomic_ld + gn;
            if (gn + 3 < p.N && (p.atomic_ld & 3) == 0) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dstg), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                           "f"(o[3]) : "memory");
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (gn + e < p.N) atomicAdd(dstg + e, o[e]);
            }
          }
        } else {
          *dst = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    if (!split) {
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(1, GEMM_EPI_THREADS);
      if (et == 0) {
        for (int c = 0; c < OUT_SLABS; ++c)
          ptx::tma_store_4d(&tmC, staging + c * (BLOCK_M * 128), n0 + OUT_COLS * c, m0, b2 * p.c_b2, b3 * p.c_b3);
        ptx::tma_store_commit();
      }
      if ((p.flags & EPI_COLSUM) && et < BLOCK_N && n0 + et < p.N) {
        // bias gradient fused into the epilogue: thread = output column, walks the 128 staged rows (conflict-free
        // under the 128B swizzle).  Rows past M hold exact zeros (zero-filled operands / mask tile).  A bf16 output
        // is summed as staged (bf16-rounded terms, fp32 sum).
        float t = 0.f;
        if constexpr (OUT16) {
          const uint8_t* slab = staging + (et >> 6) * (BLOCK_M * 128);
          const int cc = et & 63;
#pragma unroll 8
          for (int r = 0; r < BLOCK_M; ++r) {
            const uint16_t hv = *reinterpret_cast<const uint16_t*>(slab + r * 128 + ((((cc >> 3) ^ (r & 7)) << 4) | ((cc & 7) << 1)));
            t += __uint_as_float(uint32_t(hv) << 16);
          }
        } else {
          const uint8_t* slab = staging + (et >> 5) * (BLOCK_M * 128);
          const int cc = et & 31;
#pragma unroll 8
          for (int r = 0; r < BLOCK_M; ++r)
            t += *reinterpret_cast<const float*>(slab + r * 128 + ((((cc >> 2) ^ (r & 7)) << 4) | ((cc & 3) << 2)));
        }
        atomicAdd(p.colsum_out + n0 + et, t);
      }
      if (et == 0) ptx::tma_store_wait_read();   // smem may be released once the TMA engine has read it
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ persistent variant
// Same maths and epilogues as gemm_tf32_kernel, restructured as a persistent, fully decoupled pipeline: one CTA per
// SM walks a static round-robin list of output tiles; the TMA producer keeps a 4-stage operand ring (128 KB) full
// ACROSS tile boundaries, the MMA issuer alternates between two TMEM accumulators, and the epilogue warps drain one
// accumulator while the next tile's loads and MMAs are already in flight.  This keeps ~128 KB of loads in flight
// per SM at all times, which is what an HBM-bound GEMM with K = 128..512 needs (the non-persistent kernel has no
// loads in flight during its epilogue).
template <int BLOCK_N>
struct PersistLayout {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = BLOCK_N <= 64 ? 6 : 4;
  static constexpr int STAGING_BYTES = BLOCK_M * BLOCK_N * 4;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int total() { return RING_BYTES + STAGING_BYTES + 512 + BLOCK_N * 4 + 1024; }
};

constexpr int EPI_GROUPS = 2;                 // epilogue groups of four warps (one warp per TMEM lane quadrant)
constexpr int EPI_THREADS = 128 * EPI_GROUPS;
constexpr int PERSIST_THREADS = 64 + EPI_THREADS;

template <int BLOCK_N, int A_MN, int B_MN>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) gemm_tf32_persistent(const __grid_constant__ CUtensorMap tmA,
                                                                        const __grid_constant__ CUtensorMap tmB,
                                                                        const __grid_constant__ CUtensorMap tmC,
                                                                        const __grid_constant__ CUtensorMap tmAux,
                                                                        const GemmParams p, int n_tiles_n,
                                                                        int n_tiles_m, int n_z) {
  using L = PersistLayout<BLOCK_N>;
  constexpr int STAGES = L::STAGES;
  constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  constexpr int N_SLABS = BLOCK_N / 32;

  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + L::RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + L::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;        // [2] accumulator complete
  uint64_t* acc_empty = acc_full + 2;             // [2] accumulator drained by the epilogue (128 arrivals)
  uint64_t* aux_full = acc_empty + 2;             // aux tile landed in staging
  uint64_t* stage_free = aux_full + 1;            // staging free again (store has read it), 1 arrival
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_free + 1);
  float* bias_s = reinterpret_cast<float*>(staging + L::STAGING_BYTES + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = (p.flags & EPI_ATOMIC) != 0;
  const bool has_aux = (p.flags & (EPI_ADD_AUX | EPI_MASK_AUX)) != 0;
  const bool dev_rows = p.rows_dev != nullptr;     // packed rows: see gemm_tf32_kernel
  if (dev_rows && !split) n_tiles_m = min(n_tiles_m, (__ldg(p.rows_dev) + BLOCK_M - 1) / BLOCK_M);
  const int total_kb = (((dev_rows && split) ? min(p.K, __ldg(p.rows_dev)) : p.K) + BLOCK_K - 1) / BLOCK_K;
  const int kb_per = (dev_rows && split) ? (total_kb + n_z - 1) / n_z : p.kb_per_split;
  const int n_tiles = n_tiles_n * n_tiles_m * n_z;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (!split) ptx::prefetch_tmap(&tmC);
    if (has_aux) ptx::prefetch_tmap(&tmAux);
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&acc_full[a], 1); ptx::mbar_init(&acc_empty[a], EPI_THREADS); }
    ptx::mbar_init(aux_full, 1);
    ptx::mbar_init(stage_free, N_SLABS < EPI_GROUPS ? N_SLABS : EPI_GROUPS);   // one arrival per epilogue group that stores
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  arb_pdl_wait();          // everything above overlaps the previous kernel's tail; global memory is touched below
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int t, int& n0, int& m0, int& z) {
    const int nt = t % n_tiles_n;
    const int rest = t / n_tiles_n;
    n0 = nt * BLOCK_N;
    m0 = (rest % n_tiles_m) * BLOCK_M;
    z = rest / n_tiles_m;
  };
  auto k_range = [&](int z, int& kb0, int& nkb) {
    if (split) {
      kb0 = z * kb_per;
      nkb = min(total_kb, kb0 + kb_per) - kb0;
      if (nkb < 0) nkb = 0;
    } else {
      kb0 = 0;
      nkb = total_kb;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;          // running k-block counter across tiles (ring position)
      int local = 0;            // tiles processed by this CTA
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++local) {
        int n0, m0, z, kb0, nkb;
        decode(t, n0, m0, z);
        k_range(z, kb0, nkb);
        const int b2 = split ? 0 : z % p.nb2, b3 = split ? 0 : z / p.nb2;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t round = it / STAGES;
          if (round > 0) ptx::mbar_wait(&empty_bar[s], (round - 1) & 1);
          uint8_t* a_s = smem + s * L::STAGE_BYTES;
          uint8_t* b_s = a_s + A_STAGE_BYTES;
          const int k0 = (kb0 + i) * BLOCK_K;
          ptx::mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
          if (A_MN) {
            for (int c = 0; c < BLOCK_M / 32; ++c)
              ptx::tma_load_4d(a_s + c * 4096, &tmA, &full_bar[s], m0 + 32 * c, k0, b2 * p.a_b2, b3 * p.a_b3);
          } else {
            ptx::tma_load_4d(a_s, &tmA, &full_bar[s], k0, m0, b2 * p.a_b2, b3 * p.a_b3);
          }
          if (B_MN) {
            for (int c = 0; c < N_SLABS; ++c)
              ptx::tma_load_4d(b_s + c * 4096, &tmB, &full_bar[s], n0 + 32 * c, k0, b2 * p.b_b2, b3 * p.b_b3);
          } else {
            ptx::tma_load_4d(b_s, &tmB, &full_bar[s], k0, n0, b2 * p.b_b2, b3 * p.b_b3);
          }
        }
        if (has_aux) {
          // the residual / mask tile goes into the staging area, which is reused tile after tile: wait until the
          // previous tile's TMA store has read it (issued AFTER this tile's operand loads so the ring never stalls
          // behind the epilogue)
          if (local > 0) ptx::mbar_wait(stage_free, (local - 1) & 1);
          ptx::mbar_expect_tx(aux_full, L::STAGING_BYTES);
          for (int c = 0; c < N_SLABS; ++c)
            ptx::tma_load_4d(staging + c * (BLOCK_M * 128), &tmAux, aux_full, n0 + 32 * c, m0, b2 * p.c_b2, b3 * p.c_b3);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::idesc_tf32(BLOCK_M, BLOCK_N, A_MN, B_MN);
      uint32_t it = 0;
      int local = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++local) {
        int n0, m0, z, kb0, nkb;
        decode(t, n0, m0, z);
        k_range(z, kb0, nkb);
        const int acc = local & 1;
        const uint32_t use = local >> 1;      // how many times this accumulator has been used before
        if (use > 0) ptx::mbar_wait(&acc_empty[acc], (use - 1) & 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t round = it / STAGES;
          ptx::mbar_wait(&full_bar[s], round & 1);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = A_MN ? ptx::smem_desc_sw128<1>(a_addr + k * 1024, 4096, 512)
                                     : ptx::smem_desc_sw128<2>(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? ptx::smem_desc_sw128<1>(b_addr + k * 1024, 4096, 512)
                                     : ptx::smem_desc_sw128<2>(b_addr + k * 32, 16, 1024);
            ptx::mma_tf32_ss(d_tmem, da, db, idesc, (i | k) != 0 ? 1u : 0u);
          }
          ptx::mma_commit(&empty_bar[s]);
        }
        ptx::mma_commit(&acc_full[acc]);     // (with nkb == 0 this still arrives: nothing outstanding)
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // GROUPS of four warps (one warp per TMEM lane quadrant each): NGA = min(EPI_GROUPS, slabs) of them are active,
    // group g owning the 32-column slabs g, g + NGA, ... of the tile (EPI_GROUPS = 2; four groups -- one slab per group
    // on 128-wide tiles -- measured 1-3 % slower on every shape: profiles/r2/call17).  A group turns a slab TMEM -> registers -> swizzled staging and its leader
    // immediately issues that slab's TMA store; the staging slab is reclaimed lazily (cp.async.bulk.wait_group.read)
    // right before the group writes it again one tile later.  Stores, staging writes and the other groups' work
    // therefore overlap, and no barrier spans more than the 128 threads of a group.
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const int grp = (warp - 2) >> 2;                    // 0 .. EPI_GROUPS-1
    const int gt = threadIdx.x - 64 - 128 * grp;        // 0..127 inside the group
    const bool leader = gt == 0;
    constexpr int NGA = N_SLABS < EPI_GROUPS ? N_SLABS : EPI_GROUPS;   // active groups
    constexpr int SLABS_PER_GROUP = N_SLABS / NGA;
    const bool active = grp < NGA;                      // the other groups only drain the accumulator barriers
    int local = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++local) {
      int n0, m0, z, kb0, nkb;
      decode(t, n0, m0, z);
      k_range(z, kb0, nkb);
      const int b2 = split ? 0 : z % p.nb2, b3 = split ? 0 : z / p.nb2;
      const int acc = local & 1;
      const uint32_t use = local >> 1;
      if (active && (p.flags & EPI_BIAS)) {             // this group's columns only (ordered by the group barriers)
        for (int j = gt; j < 32 * SLABS_PER_GROUP; j += 128) {
          const int col = 32 * (grp + NGA * (j >> 5)) + (j & 31);
          bias_s[col] = (n0 + col < p.N) ? p.bias[n0 + col] : 0.0f;
        }
      }
      // EPI_MASK_BITS: this row's mask words of the group's slabs, fetched before the wait for the accumulator
      uint32_t mword[2] = {0u, 0u};
      if (active && (p.flags & EPI_MASK_BITS) && m0 + row < p.M) {
#pragma unroll
        for (int ci = 0; ci < (SLABS_PER_GROUP < 2 ? SLABS_PER_GROUP : 2); ++ci) {
          const int c = grp + NGA * ci;
          if (n0 + 32 * c < p.N) mword[ci] = p.bits[(long long)(m0 + row) * (p.N >> 5) + ((n0 >> 5) + c)];
        }
      }
      ptx::mbar_wait(&acc_full[acc], use & 1);
      ptx::tc_fence_after();
      if (has_aux) ptx::mbar_wait(aux_full, local & 1);
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      if (active) {
#pragma unroll 1
        for (int ci = 0; ci < SLABS_PER_GROUP; ++ci) {
          const int c = grp + NGA * ci;
          // reclaim the slab: every store this leader committed except the most recent (SLABS_PER_GROUP - 1) ones has
          // been read out of shared memory -- in particular the one that used slab c a tile ago.  With an aux tile
          // the leader already drained its stores before the producer refilled the staging area.
          if (!split && !has_aux && leader && local > 0) {
            if constexpr (SLABS_PER_GROUP == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else ptx::tma_store_wait_read();
          }
          ptx::named_bar_sync(1 + grp, 128);
          uint32_t v[32];
          if (nkb > 0) {
            ptx::tmem_ld_32x32(d_tmem + (uint32_t(32 * q) << 16) + uint32_t(32 * c), v);
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          uint8_t* slab_row = staging + c * (BLOCK_M * 128) + row * 128;
          uint32_t* bits_at = (p.bits && m0 + row < p.M && n0 + 32 * c < p.N)
                                  ? p.bits + (long long)(m0 + row) * (p.N >> 5) + ((n0 >> 5) + c) : nullptr;
          const uint32_t word_in = ci == 0 ? mword[0] : mword[1];
          if (!epi_chunk_dispatch(p.flags, v, slab_row, row, bias_s + 32 * c, p.alpha, bits_at, word_in)) {
#pragma unroll
          for (int piece = 0; piece < 8; ++piece) {
            float4* dst = reinterpret_cast<float4*>(slab_row + ((piece ^ (row & 7)) << 4));
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = piece * 4 + e;
              float x = __uint_as_float(v[j]) * p.alpha;
              if (p.flags & EPI_BIAS) x += bias_s[32 * c + j];
              if (p.flags & EPI_RELU) x = fmaxf(x, 0.0f);
              if (p.flags & EPI_DROPOUT) {
                const unsigned long long idx = (unsigned long long)(m0 + row) * (unsigned long long)p.N + (n0 + 32 * c + j);
                x = drop_keep(idx, p.drop.seed, p.drop.thresh) ? x * p.drop.scale : 0.0f;
              }
              o[e] = x;
            }
            if (has_aux) {
              const float4 a = *dst;
              if (p.flags & EPI_ADD_AUX) { o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w; }
              if (p.flags & EPI_MASK_AUX) {
                o[0] = a.x > 0.f ? o[0] : 0.f; o[1] = a.y > 0.f ? o[1] : 0.f;
                o[2] = a.z > 0.f ? o[2] : 0.f; o[3] = a.w > 0.f ? o[3] : 0.f;
              }
            }
            if (split) {
              const int gm = m0 + row, gn = n0 + 32 * c + piece * 4;
              if (gm < p.M && gn < p.N) {
                float* dstg = p.atomic_out + (long long)gm * p.atomic_ld + gn;
                if (gn + 3 < p.N && (p.atomic_ld & 3) == 0) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dstg), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                               "f"(o[3]) : "memory");
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (gn + e < p.N) atomicAdd(dstg + e, o[e]);
                }
              }
            } else {
              *dst = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
          }
          if (!split) {
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(1 + grp, 128);
            if (leader) {
              ptx::tma_store_4d(&tmC, staging + c * (BLOCK_M * 128), n0 + 32 * c, m0, b2 * p.c_b2, b3 * p.c_b3);
              ptx::tma_store_commit();
            }
            if (p.flags & EPI_COLSUM) {
              // bias gradient fused into the epilogue: 4 threads per output column, 32 staged rows each (conflict-free
              // under the 128B swizzle), combined by shuffles.  Rows past M hold exact zeros.  The slab is not
              // rewritten before this group's next barrier.
              const int cc = gt >> 2, seg = gt & 3;
              const uint8_t* slab = staging + c * (BLOCK_M * 128);
              float tsum = 0.f;
#pragma unroll 8
              for (int r = 32 * seg; r < 32 * seg + 32; ++r)
                tsum += *reinterpret_cast<const float*>(slab + r * 128 + ((((cc >> 2) ^ (r & 7)) << 4) | ((cc & 3) << 2)));
              tsum += __shfl_xor_sync(0xffffffffu, tsum, 1);
              tsum += __shfl_xor_sync(0xffffffffu, tsum, 2);
              if (seg == 0 && n0 + 32 * c + cc < p.N) atomicAdd(p.colsum_out + n0 + 32 * c + cc, tsum);
            }
          }
        }
      }
      // this accumulator may be overwritten by the MMA warp from now on
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acc_empty[acc]);
      if (!split && has_aux && active && leader) {     // the producer refills the staging area with the next aux tile
        ptx::tma_store_wait_read();
        ptx::mbar_arrive(stage_free);
      }
    }
    if (!split && active && leader) ptx::tma_store_wait_read();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// MMA operands are loaded through TFLOAT32 tensor maps: the TMA unit rounds fp32 -> tf32 (instead of the tensor
// core truncating the mantissa), which removes the systematic bias truncation puts on gradients.
static int g_round_on_load = 1;
void set_tf32_round_on_load(int enable) { g_round_on_load = enable; }

int make_tmap_4d(void* out, const TRef& t, TmapBox box, int atom32, int as_tf32) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { arb_set_error("cuTensorMapEncodeTiled is not available from this driver"); return ARB_E_CUDA; }
  const int64_t esz = t.bf16 ? 2 : 4;
  const int64_t per16 = 16 / esz;                 // elements per 16 bytes
  cuuint64_t gdim[4], gstride[3];
  cuuint32_t bx[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { gdim[i] = cuuint64_t(t.dim[i] > 0 ? t.dim[i] : 1); bx[i] = box.b[i]; }
  for (int i = 1; i < 4; ++i) {
    // a broadcast / unused dimension (extent 1) still needs a legal (multiple of 16 B) stride
    int64_t s = t.stride[i];
    if (gdim[i] == 1 && (s <= 0 || (s * esz) % 16 != 0))
      s = int64_t(gdim[0]) * esz >= 16 ? ((int64_t(gdim[0]) + per16 - 1) / per16) * per16 : per16;
    gstride[i - 1] = cuuint64_t(s) * esz;
    if (gstride[i - 1] % 16 != 0) { arb_set_error("tensor map: strides must be multiples of 16 bytes"); return ARB_E_INVALID_ARG; }
  }
  if ((reinterpret_cast<uintptr_t>(t.ptr) & 15) != 0) { arb_set_error("tensor map: base must be 16-byte aligned"); return ARB_E_INVALID_ARG; }
  if (t.bf16 && atom32 == 1) { arb_set_error("tensor map: the 32-byte-atom swizzle is a tf32 layout"); return ARB_E_INVALID_ARG; }
  const CUtensorMapDataType dt = t.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                        : ((as_tf32 && g_round_on_load) ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                                                                        : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), dt, 4, const_cast<void*>(t.ptr),
                   gdim, gstride, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE
                               : (atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[256];
    std::snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d): dim=(%llu,%llu,%llu,%llu) box=(%u,%u,%u,%u)", int(r),
                  (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
                  (unsigned long long)gdim[3], bx[0], bx[1], bx[2], bx[3]);
    arb_set_error(msg);
    return ARB_E_CUDA;
  }
  return ARB_OK;
}

// launch name for the per-kernel profile table: shape, operand layouts and what the epilogue does
static void gemm_prof_name(char (&out)[56], const GemmDesc& d, const char* variant) {
  const char* kind = (d.flags & EPI_ATOMIC) ? "wgrad" : (d.nb2 * d.nb3 > 1 ? "batched" : (d.b_mn ? "dgrad" : "fwd"));
  std::snprintf(out, sizeof out, "gemm_%s%s[%s M%d N%d K%d%s%s%s]", variant, d.A.bf16 ? (d.C.bf16 ? "_bf16o" : "_bf16") : "",
                kind, d.M, d.N, d.K, (d.flags & EPI_RELU) ? " relu" : "", (d.flags & EPI_ADD_AUX) ? " +res" : "",
                (d.flags & EPI_MASK_AUX) ? " mask" : "");
}

// accounting only: a launch over packed rows (rows_dev) processes arb_row_frac() of its nominal M (or K, split-K)
static double live_m(const GemmDesc& d) { return double(d.M) * ((d.rows_dev && !(d.flags & EPI_ATOMIC)) ? arb_row_frac() : 1.0); }
static double live_k(const GemmDesc& d) { return double(d.K) * ((d.rows_dev && (d.flags & EPI_ATOMIC)) ? arb_row_frac() : 1.0); }
static int g_persistent = ARB_DEFAULT_GEMM_PERSISTENT;   // 0: never, 1: wherever supported, 2: auto
void set_gemm_persistent(int on) { g_persistent = on; }

template <int BLOCK_N, int A_MN, int B_MN>
static int launch_persistent_t(const GemmDesc& d, const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC,
                               const CUtensorMap& tX, const GemmParams& p, dim3 tiles, cudaStream_t st) {
  auto kern = gemm_tf32_persistent<BLOCK_N, A_MN, B_MN>;
  constexpr int smem = PersistLayout<BLOCK_N>::total();
  static bool configured[ARB_MAX_DEVICES] = {};
  static int n_sm_of[ARB_MAX_DEVICES] = {};
  const int dev_slot = arb_device_slot();
  if (!configured[dev_slot]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      arb_set_error("gemm_tf32 (persistent): cannot raise the dynamic shared memory limit");
      return ARB_E_CUDA;
    }
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    n_sm_of[dev_slot] = n;
    configured[dev_slot] = true;
  }
  const int n_sm = n_sm_of[dev_slot];
  const long long n_tiles = (long long)tiles.x * tiles.y * tiles.z;
  const int grid = int(std::min<long long>(n_tiles, n_sm));
  {
    const double nb = double(d.nb2) * double(d.nb3);
    const double has_x = (d.flags & (EPI_ADD_AUX | EPI_MASK_AUX)) ? 1.0 : 0.0;
    char pname[56];
    gemm_prof_name(pname, d, "persist");
    const double dM = live_m(d), dK = live_k(d);
    ProfScope ps(ARB_PROF_GEMM, 2.0 * dM * double(d.N) * dK * nb, st,
                 4.0 * nb * (dM * dK + double(d.N) * dK + (1.0 + has_x) * dM * d.N), pname);
    arb_launch(kern, dim3(grid), dim3(PERSIST_THREADS), smem, st, tA, tB, tC, tX, p, int(tiles.x), int(tiles.y), int(tiles.z));
  }
  arb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { arb_set_error(cudaGetErrorString(e)); return ARB_E_CUDA; }
  return ARB_OK;
}

// bf16 operands (kind::f16): the one-tile-per-CTA kernel; 4-stage ring for the split-K weight gradients, else 3 stages
// (a k-block is 64 elements, so K = 256 is four blocks); fp32 or bf16 output.
template <int BLOCK_N, int A_MN, int B_MN>
static int launch_bf16_t(const GemmDesc& d, const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC,
                         const CUtensorMap& tX, const GemmParams& p, dim3 grid, cudaStream_t st) {
  constexpr bool WGRAD = (A_MN == 1 && B_MN == 1);
  const bool drop = (p.flags & EPI_DROPOUT) != 0;
  if (drop && (A_MN || B_MN)) { arb_set_error("gemm_bf16: dropout epilogue needs K-major operands"); return ARB_E_UNSUPPORTED; }
  void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, GemmParams) = nullptr;
  int smem = 0, slot = 0;
  if constexpr (WGRAD) {
    kern = gemm_tf32_kernel<BLOCK_N, 1, 1, false, 4, true, false>; smem = SmemLayout<BLOCK_N, 4>::total(); slot = 0;
  } else if constexpr (BLOCK_N >= 64) {
    constexpr bool CAN_DROP = (A_MN == 0 && B_MN == 0);
    if (d.C.bf16) {
      if (drop) { kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, CAN_DROP, 3, true, true>; slot = 1; }
      else      { kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, false, 3, true, true>; slot = 2; }
    } else {
      if (drop) { kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, CAN_DROP, 3, true, false>; slot = 3; }
      else      { kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, false, 3, true, false>; slot = 4; }
    }
    smem = SmemLayout<BLOCK_N, 3>::total();
  }
  if (!kern) { arb_set_error("gemm_bf16: block_n must be 64 or 128"); return ARB_E_UNSUPPORTED; }
  static bool configured[ARB_MAX_DEVICES][5] = {};
  const int dev_slot = arb_device_slot();
  if (!configured[dev_slot][slot]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      arb_set_error("gemm_bf16: cannot raise the dynamic shared memory limit");
      return ARB_E_CUDA;
    }
    configured[dev_slot][slot] = true;
  }
  {
    const double nb = double(d.nb2) * double(d.nb3);
    const double has_x = (d.flags & (EPI_ADD_AUX | EPI_MASK_AUX)) ? 1.0 : 0.0;
    const double osz = d.C.bf16 ? 2.0 : 4.0;
    char pname[56];
    gemm_prof_name(pname, d, "tile");
    const double dM = live_m(d), dK = live_k(d);
    ProfScope ps(ARB_PROF_GEMM, 2.0 * dM * double(d.N) * dK * nb, st,
                 nb * (2.0 * (dM * dK + double(d.N) * dK) + osz * (1.0 + has_x) * dM * d.N), pname);
    arb_launch(kern, grid, dim3(GEMM_THREADS), smem, st, tA, tB, tC, tX, p);
  }
  arb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { arb_set_error(cudaGetErrorString(e)); return ARB_E_CUDA; }
  return ARB_OK;
}

template <int BLOCK_N, int A_MN, int B_MN>
static int launch_t(const GemmDesc& d, const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC,
                    const CUtensorMap& tX, const GemmParams& p, dim3 grid, cudaStream_t st) {
  if (d.A.bf16) return launch_bf16_t<BLOCK_N, A_MN, B_MN>(d, tA, tB, tC, tX, p, grid, st);
  // Mode 2 (default), measured per launch in round 2 after the epilogue was specialised (profiles/r2): the persistent
  // pipeline wins on every unbatched, non-split shape (short-K wide-N forward linears 0.76-0.86 of the HBM roof against
  // 0.58-0.62 for one-tile CTAs; K >= 256: 0.93-0.98) EXCEPT short-K products with a residual / mask tile, whose aux
  // load it can only issue once the previous tile's stores have left the staging area (dgrad N512 K128 mask: 0.70
  // against 0.85; the K128 +res projection ties).  (Giving the aux tile a shared-memory buffer of its own, at the price
  // of one ring stage, measured slower on every shape: 19.17 against 18.63 ms per cfg2 step -- not kept.)
  const bool has_aux_tile = (p.flags & (EPI_ADD_AUX | EPI_MASK_AUX)) != 0;
  // Mode 3 (measurement): as 2, but the short-K products with an aux tile go to the persistent kernel too (under packed
  // rows half of a one-tile grid are CTAs of dead tiles that only exit; the persistent kernel has none).
  const bool pick = !(A_MN == 1 && B_MN == 1) && !(p.flags & EPI_ATOMIC) && d.nb2 == 1 && d.nb3 == 1 &&
                    (d.K >= 256 || !has_aux_tile || g_persistent == 3);
  if (g_persistent == 1 || (g_persistent >= 2 && pick))
    return launch_persistent_t<BLOCK_N, A_MN, B_MN>(d, tA, tB, tC, tX, p, grid, st);
  // dropout epilogue is a separate instantiation (forward linears only) so the common path carries no mask code;
  // ring depth: 4 stages for the long split-K loops of the weight gradients (1 CTA/SM), 3 for K >= 256
  // (2 CTAs/SM), 2 for the short contractions (3-4 CTAs/SM)
  constexpr bool CAN_DROP = (A_MN == 0 && B_MN == 0);
  constexpr bool WGRAD = (A_MN == 1 && B_MN == 1);
  const bool drop = (p.flags & EPI_DROPOUT) != 0;
  if (drop && !CAN_DROP) { arb_set_error("gemm_tf32: dropout epilogue needs K-major operands"); return ARB_E_UNSUPPORTED; }
  const bool deep = !WGRAD && d.K >= 256;
  void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, GemmParams);
  int smem, slot;
  if (WGRAD) {
    kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, false, 4>; smem = SmemLayout<BLOCK_N, 4>::total(); slot = 0;
  } else if (drop) {
    if (deep) { kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, CAN_DROP, 3>; smem = SmemLayout<BLOCK_N, 3>::total(); slot = 1; }
    else      { kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, CAN_DROP, 2>; smem = SmemLayout<BLOCK_N, 2>::total(); slot = 2; }
  } else {
    if (deep) { kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, false, 3>; smem = SmemLayout<BLOCK_N, 3>::total(); slot = 3; }
    else      { kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, false, 2>; smem = SmemLayout<BLOCK_N, 2>::total(); slot = 4; }
  }
  static bool configured[ARB_MAX_DEVICES][5] = {};
  const int dev_slot = arb_device_slot();
  if (!configured[dev_slot][slot]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      arb_set_error("gemm_tf32: cannot raise the dynamic shared memory limit");
      return ARB_E_CUDA;
    }
    configured[dev_slot][slot] = true;
  }
  {
    const double nb = double(d.nb2) * double(d.nb3);
    const double has_x = (d.flags & (EPI_ADD_AUX | EPI_MASK_AUX)) ? 1.0 : 0.0;
    char pname[56];
    gemm_prof_name(pname, d, "tile");
    const double dM = live_m(d), dK = live_k(d);
    ProfScope ps(ARB_PROF_GEMM, 2.0 * dM * double(d.N) * dK * nb, st,
                 4.0 * nb * (dM * dK + double(d.N) * dK + (1.0 + has_x) * dM * d.N), pname);
    arb_launch(kern, grid, dim3(GEMM_THREADS), smem, st, tA, tB, tC, tX, p);
  }
  arb_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { arb_set_error(cudaGetErrorString(e)); return ARB_E_CUDA; }
  return ARB_OK;
}

int launch_gemm_tf32(const GemmDesc& d, cudaStream_t st) {
  if (d.M <= 0 || d.N <= 0 || d.K < 0) { arb_set_error("gemm_tf32: bad shape"); return ARB_E_INVALID_ARG; }
  if (d.block_n != 32 && d.block_n != 64 && d.block_n != 128) { arb_set_error("gemm_tf32: block_n must be 32/64/128"); return ARB_E_INVALID_ARG; }
  const bool split = (d.flags & EPI_ATOMIC) != 0;
  if (split && (d.nb2 != 1 || d.nb3 != 1 || !d.atomic_out)) { arb_set_error("gemm_tf32: split-K needs an unbatched problem and atomic_out"); return ARB_E_INVALID_ARG; }
  if (!split && d.split_k != 1) { arb_set_error("gemm_tf32: split_k > 1 needs EPI_ATOMIC"); return ARB_E_INVALID_ARG; }
  const bool in16 = d.A.bf16 != 0, out16 = d.C.bf16 != 0;
  if ((d.B.bf16 != 0) != in16) { arb_set_error("gemm: A and B must have the same element type"); return ARB_E_INVALID_ARG; }
  if (out16 && (!in16 || split || d.block_n < 64)) { arb_set_error("gemm: a bf16 output needs bf16 operands, no split-K and block_n >= 64"); return ARB_E_INVALID_ARG; }
  if ((d.flags & (EPI_ADD_AUX | EPI_MASK_AUX)) && (d.Aux.bf16 != 0) != out16) { arb_set_error("gemm: the aux tile must have the output's element type"); return ARB_E_INVALID_ARG; }
  if (in16 && d.block_n < 64) { arb_set_error("gemm: bf16 operands need block_n >= 64 (a 128-byte row holds 64 elements)"); return ARB_E_INVALID_ARG; }
  if (in16 && (d.nb2 != 1 || d.nb3 != 1)) { arb_set_error("gemm: bf16 operands serve the unbatched linears only"); return ARB_E_UNSUPPORTED; }
  const uint32_t kel = in16 ? 64u : 32u;          // K elements per 128-byte row
  const uint32_t ocol = out16 ? 64u : 32u;        // output columns per 128-byte staging row
  alignas(64) CUtensorMap tA, tB, tC, tX;
  int rc;
  if ((rc = make_tmap_4d(&tA, d.A, d.a_mn ? TmapBox{{kel, kel, 1, 1}} : TmapBox{{kel, 128, 1, 1}}, d.a_mn && !in16, 1))) return rc;
  if ((rc = make_tmap_4d(&tB, d.B, d.b_mn ? TmapBox{{kel, kel, 1, 1}} : TmapBox{{kel, uint32_t(d.block_n), 1, 1}}, d.b_mn && !in16, 1))) return rc;
  if (!split) {
    if ((rc = make_tmap_4d(&tC, d.C, TmapBox{{ocol, 128, 1, 1}}, 0, 0))) return rc;
  } else {
    tC = tA;
  }
  if (d.flags & (EPI_ADD_AUX | EPI_MASK_AUX)) {
    if ((rc = make_tmap_4d(&tX, d.Aux, TmapBox{{ocol, 128, 1, 1}}, 0, 0))) return rc;
  } else {
    tX = tA;
  }
  const int total_kb = (d.K + int(kel) - 1) / int(kel);
  const int splits = split ? std::max(1, std::min(d.split_k, total_kb)) : 1;
  GemmParams p;
  p.M = d.M; p.N = d.N; p.K = d.K; p.nb2 = d.nb2;
  p.a_b2 = d.a_b2; p.a_b3 = d.a_b3; p.b_b2 = d.b_b2; p.b_b3 = d.b_b3; p.c_b2 = d.c_b2; p.c_b3 = d.c_b3;
  p.flags = d.flags; p.alpha = d.alpha; p.bias = d.bias; p.atomic_out = d.atomic_out; p.atomic_ld = d.atomic_ld;
  p.drop = d.drop;
  p.colsum_out = d.colsum_out;
  p.rows_dev = d.rows_dev;
  p.bits = d.bits;
  if (d.flags & (EPI_RELU_BITS | EPI_MASK_BITS)) {
    if (!d.bits || d.N % 32 || out16 || split || d.nb2 != 1 || d.nb3 != 1 || (d.flags & (EPI_DROPOUT | EPI_ADD_AUX | EPI_MASK_AUX)) ||
        ((d.flags & EPI_RELU_BITS) && !(d.flags & EPI_RELU))) {
      arb_set_error("gemm: the ReLU bit mask needs an unbatched fp32 output with N % 32 == 0, no dropout and no aux tile");
      return ARB_E_INVALID_ARG;
    }
    const int core = d.flags & ~EPI_COLSUM;
    if (core != (EPI_BIAS | EPI_RELU | EPI_RELU_BITS) && core != EPI_MASK_BITS && !((core == (EPI_RELU | EPI_RELU_BITS)) && d.bias)) {
      arb_set_error("gemm: the ReLU bit mask serves bias + ReLU (forward) and the plain mask (backward) only");
      return ARB_E_INVALID_ARG;
    }
  }
  if (d.rows_dev && (d.nb2 != 1 || d.nb3 != 1)) { arb_set_error("gemm: a device-side row count serves unbatched launches only"); return ARB_E_INVALID_ARG; }
  if ((d.flags & EPI_COLSUM) && (!d.colsum_out || split)) { arb_set_error("gemm_tf32: EPI_COLSUM needs colsum_out and a non-split launch"); return ARB_E_INVALID_ARG; }
  if ((d.flags & EPI_DROPOUT) && d.drop.thresh == 0) p.flags &= ~EPI_DROPOUT;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  const int eff_splits = split ? (total_kb + p.kb_per_split - 1) / std::max(1, p.kb_per_split) : 1;
  dim3 grid((d.N + d.block_n - 1) / d.block_n, (d.M + BLOCK_M - 1) / BLOCK_M, split ? std::max(1, eff_splits) : d.nb2 * d.nb3);
#define ARB_GEMM_CASE(BN, AM, BM) \
  if (d.block_n == BN && d.a_mn == AM && d.b_mn == BM) return launch_t<BN, AM, BM>(d, tA, tB, tC, tX, p, grid, st);
  ARB_GEMM_CASE(32, 0, 0) ARB_GEMM_CASE(32, 0, 1) ARB_GEMM_CASE(32, 1, 0) ARB_GEMM_CASE(32, 1, 1)
  ARB_GEMM_CASE(64, 0, 0) ARB_GEMM_CASE(64, 0, 1) ARB_GEMM_CASE(64, 1, 0) ARB_GEMM_CASE(64, 1, 1)
  ARB_GEMM_CASE(128, 0, 0) ARB_GEMM_CASE(128, 0, 1) ARB_GEMM_CASE(128, 1, 0) ARB_GEMM_CASE(128, 1, 1)
#undef ARB_GEMM_CASE
  arb_set_error("gemm_tf32: unsupported configuration");
  return ARB_E_UNSUPPORTED;
}

}  // namespace arb

// ------------------------------------------------------------------------------------------------ C ABI (unit-test / building-block entry)
// C[b][M,N] = epilogue(alpha * A[b] op B[b]) on plain row-major fp32 matrices.
//   a_mn = 0: A is [M,K] row-major;  a_mn = 1: A is stored transposed, [K,M] row-major.
//   b_mn = 0: B is [N,K] row-major (an nn.Linear weight);  b_mn = 1: B is [K,N] row-major.
//   batch > 1: operands are `batch` consecutive matrices (stride = rows*cols), unless the stride argument is 0.
extern "C" void arb_set_tf32_round_on_load(int32_t enable) { arb::set_tf32_round_on_load(enable); }
extern "C" void arb_set_gemm_persistent(int32_t on) { arb::set_gemm_persistent(on); }

extern "C" int32_t arb_gemm_tf32(const float* A, const float* B, float* C, const float* aux, const float* bias,
                                 int32_t M, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn, int32_t batch,
                                 int64_t a_bstride, int64_t b_bstride, int64_t c_bstride, int32_t block_n,
                                 int32_t flags, float alpha, int32_t split_k, void* stream) {
  using namespace arb;
  if (!A || !B || !C) { arb_set_error("arb_gemm_tf32: null pointer"); return ARB_E_INVALID_ARG; }
  GemmDesc d;
  d.M = M; d.N = N; d.K = K; d.a_mn = a_mn; d.b_mn = b_mn; d.block_n = block_n; d.flags = flags; d.alpha = alpha;
  d.bias = bias; d.nb2 = batch; d.nb3 = 1; d.split_k = split_k;
  d.a_b2 = a_bstride != 0; d.b_b2 = b_bstride != 0; d.c_b2 = c_bstride != 0;
  auto mat = [](const float* p, int64_t inner, int64_t rows, int64_t nb, int64_t bstride) {
    TRef t; t.ptr = p; t.dim[0] = inner; t.dim[1] = rows; t.dim[2] = bstride ? nb : 1; t.dim[3] = 1;
    t.stride[0] = 1; t.stride[1] = inner; t.stride[2] = bstride; t.stride[3] = 0; return t;
  };
  d.A = a_mn ? mat(A, M, K, batch, a_bstride) : mat(A, K, M, batch, a_bstride);
  d.B = b_mn ? mat(B, N, K, batch, b_bstride) : mat(B, K, N, batch, b_bstride);
  d.C = mat(C, N, M, batch, c_bstride);
  if (aux) d.Aux = mat(aux, N, M, batch, c_bstride);
  if (flags & EPI_ATOMIC) { d.atomic_out = C; d.atomic_ld = N; }
  return launch_gemm_tf32(d, static_cast<cudaStream_t>(stream));
}

// The same building block with bf16 operands (tcgen05 kind::f16, fp32 accumulation): A, B are bfloat16 matrices in the
// layouts described above; C (and aux, if given) is bfloat16 when out_bf16 != 0, else fp32.  With EPI_ATOMIC
// (split-K) C must be fp32 and is accumulated into.
extern "C" int32_t arb_gemm_bf16(const void* A, const void* B, void* C, const void* aux, const float* bias, int32_t M,
                                 int32_t N, int32_t K, int32_t a_mn, int32_t b_mn, int32_t block_n, int32_t flags,
                                 float alpha, int32_t split_k, int32_t out_bf16, float* colsum_out, void* stream) {
  using namespace arb;
  if (!A || !B || !C) { arb_set_error("arb_gemm_bf16: null pointer"); return ARB_E_INVALID_ARG; }
  GemmDesc d;
  d.M = M; d.N = N; d.K = K; d.a_mn = a_mn; d.b_mn = b_mn; d.block_n = block_n; d.flags = flags; d.alpha = alpha;
  d.bias = bias; d.split_k = split_k; d.colsum_out = colsum_out;
  auto mat = [](const void* p, int64_t inner, int64_t rows, int is16) {
    TRef t; t.ptr = p; t.dim[0] = inner; t.dim[1] = rows; t.stride[0] = 1; t.stride[1] = inner; t.bf16 = is16; return t;
  };
  d.A = a_mn ? mat(A, M, K, 1) : mat(A, K, M, 1);
  d.B = b_mn ? mat(B, N, K, 1) : mat(B, K, N, 1);
  d.C = mat(C, N, M, out_bf16 != 0);
  if (aux) d.Aux = mat(aux, N, M, out_bf16 != 0);
  if (flags & EPI_ATOMIC) { d.atomic_out = static_cast<float*>(C); d.atomic_ld = N; }
  return launch_gemm_tf32(d, static_cast<cudaStream_t>(stream));
}
