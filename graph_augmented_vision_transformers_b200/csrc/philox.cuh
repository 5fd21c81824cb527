// philox.cuh - counter-based keep-mask generation shared by the dropout edges (edges.cu, embed.cu).
#pragma once
#include "common.cuh"

namespace gvit {

// ---- Philox-4x32-R (Salmon et al. 2011), counter = (offset + i/8), key = seed ---------------------
// The keep masks use R = 7: the fewest rounds that are Crush-resistant in the paper (BigCrush-clean); the customary 10
// add safety margin that a dropout mask does not need, and the rounds are a third of the mask kernels' integer work.
constexpr int GVIT_PHILOX_ROUNDS = 7;
template <int R>
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// 8 keep decisions from ONE Philox block (16 random bits per element): bit j of the result is 1 when element j is
// kept.  P(keep) = 1 - thresh16 / 65536 with thresh16 = round(p * 65536), i.e. p is honoured to 1.5e-5.
__device__ __forceinline__ uint32_t keep_bits8(uint64_t seed, uint64_t counter, uint32_t thresh16) {
  const uint4 r = philox4x32<GVIT_PHILOX_ROUNDS>(make_uint4((uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) bits |= (((w[j >> 1] >> (16 * (j & 1))) & 0xffffu) >= thresh16 ? 1u : 0u) << j;
  return bits;
}
__device__ __forceinline__ uint32_t dropout_thresh16(float p) { return (uint32_t)__float2int_rn(p * 65536.0f); }

}  // namespace gvit
