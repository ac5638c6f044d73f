// philox.cuh — Philox4x32-10 counter-based generator and the N(0,1) draws built on it (shared by the VP-SDE step
// kernel and the latent prior's initial draw).  Restated in numpy by oracle/philox_ref.py.
#pragma once
#include "common.cuh"

namespace tcs {

// ---- Philox4x32-10 (Salmon et al. 2011), key = seed, counter = (pixel group, stream word, gidx lo, gidx hi)
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ float u01(uint32_t r) { return (static_cast<float>(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// four N(0,1) draws for pixels 4*group .. 4*group+3 of sample gidx, stream word `word`
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long gidx, uint32_t word,
                                                 uint32_t group) {
  const Philox4 r = philox4x32_10(group, word, static_cast<uint32_t>(gidx), static_cast<uint32_t>(gidx >> 32),
                                  static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  // Box-Muller on the SFU: lg2.approx / sin.approx / cos.approx (abs. error ~1e-6 on N(0,1) draws), so that the
  // update kernel stays HBM-bound instead of being bound by libm's logf / sincospif
  const float r0 = sqrtf(-1.3862943611198906f * __log2f(u01(r.x)));   // sqrt(-2 ln u) = sqrt(-2 ln2 log2 u)
  const float r1 = sqrtf(-1.3862943611198906f * __log2f(u01(r.z)));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u01(r.y) - 3.141592653589793f, &s0, &c0);   // angle in (-pi, pi): best SFU accuracy
  __sincosf(6.283185307179586f * u01(r.w) - 3.141592653589793f, &s1, &c1);
  return make_float4(-r0 * c0, -r0 * s0, -r1 * c1, -r1 * s1);              // cos(a - pi) = -cos a, sin(a - pi) = -sin a
}

}  // namespace tcs
