// Counter-based dropout masks: keep(element index, site seed) is a pure function, so the backward pass regenerates
// the forward's mask instead of storing it.  (The reference uses torch's Philox stream, transformer.py:105,155,227;
// a fused kernel cannot reproduce that stream, so parity under dropout is statistical -- SURVEY.md section 7.)
#pragma once
#include <cstdint>

namespace arb {

struct DropSite {
  uint32_t seed;      // per (call, layer, site) seed
  uint32_t thresh;    // drop iff hash < thresh  (thresh = p * 2^32); 0 disables the site
  float scale;        // 1 / (1 - p)
};

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ bool drop_keep(unsigned long long idx, uint32_t seed, uint32_t thresh) {
  if (thresh == 0) return true;   // disabled site (may still carry a scale)
  uint32_t h = mix32(uint32_t(idx) ^ seed);
  h = mix32(h + uint32_t(idx >> 32) * 0x9e3779b1u + 0x7f4a7c15u);
  return h >= thresh;
}

inline DropSite make_drop_site(uint64_t call_seed, int layer, int site, float p) {
  DropSite d{0u, 0u, 1.0f};
  if (p <= 0.0f) return d;
  uint64_t z = call_seed + 0x9e3779b97f4a7c15ull * uint64_t(layer * 8 + site + 1);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z ^= z >> 31;
  d.seed = uint32_t(z) ^ uint32_t(z >> 32);
  double t = double(p) * 4294967296.0;
  d.thresh = t >= 4294967295.0 ? 0xffffffffu : uint32_t(t);
  if (d.thresh == 0) d.thresh = 1;
  d.scale = 1.0f / (1.0f - p);
  return d;
}

enum { SITE_FC = 0, SITE_ATTN_P = 1, SITE_ATTN_OUT = 2, SITE_FFN_HID = 3, SITE_FFN_OUT = 4 };

}  // namespace arb
