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

// single-instruction SFU forms (the arguments are never denormal: u01 >= 2^-25)
__device__ __forceinline__ float sfu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sfu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sfu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sfu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// four N(0,1) draws for pixels 4*group .. 4*group+3 of sample gidx, stream word `word`
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long gidx, uint32_t word,
                                                 uint32_t group) {
  const Philox4 r = philox4x32_10(group, word, static_cast<uint32_t>(gidx), static_cast<uint32_t>(gidx >> 32),
                                  static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  // Box-Muller on the SFU: lg2 / sqrt / sin / cos .approx, one MUFU each (abs. error ~1e-6 on N(0,1) draws).  The update
  // kernel is bound by instruction issue, not by HBM, as long as this costs more than a few dozen instructions
  // (ncu, round 2: 272 instructions per warp with libm-style sqrtf / __sincosf and IEEE divisions, 75 % issue-active).
  const float r0 = sfu_sqrt(-1.3862943611198906f * sfu_lg2(u01(r.x)));   // sqrt(-2 ln u) = sqrt(-2 ln2 log2 u)
  const float r1 = sfu_sqrt(-1.3862943611198906f * sfu_lg2(u01(r.z)));
  const float a0 = 6.283185307179586f * u01(r.y) - 3.141592653589793f;     // angle in (-pi, pi): best SFU accuracy
  const float a1 = 6.283185307179586f * u01(r.w) - 3.141592653589793f;
  return make_float4(-r0 * sfu_cos(a0), -r0 * sfu_sin(a0), -r1 * sfu_cos(a1), -r1 * sfu_sin(a1));   // cos(a - pi) = -cos a
}

}  // namespace tcs
