// umma_attn_probe.cu — hardware probe (not part of libtcs) for the building blocks of the fused attention-block kernel:
//   (1) tcgen05.mma with the A operand in TENSOR MEMORY (bf16 packed two per 32-bit column, written with tcgen05.st),
//   (2) an MN-major SWIZZLE_128B B operand whose N extent (48) is a partial 64-element atom  (O = P V, V rows = keys),
//   (3) cp.async.bulk (1-D, no tensor map) into shared memory completing on an mbarrier,
//   (4) operand rows written into a PEER CTA's shared memory (DSMEM, generic proxy) and read by that CTA's MMAs.
// One cluster of two CTAs: CTA 0 computes S = Q K^T (128 x 256, K = 48) and O = P V (128 x 48, K = 256) for one head;
// CTA 1 only delivers key rows 128..255 of K and V into CTA 0's shared memory.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o build/umma_attn_probe tools/umma_attn_probe.cu
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../vae-diffusion-toy-crystals_b200/csrc/tc_ptx.cuh"

namespace cg = cooperative_groups;
using namespace tcs;

struct Args {
  int a_mode;   // 0: A (Q, P) in tensor memory; 1: A in shared memory (K-major SWIZZLE_128B)
  int lbo, sbo; // descriptor fields (bytes) of the MN-major V operand
  int vmajor;   // 1: b_major = MN for the PV product
};

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

constexpr int QCOL = 400, OCOL = 424;
// shared memory (1024-aligned): K [256][128 B] | V [256][128 B] | Qs [128][128 B] | Ps [4][128][128 B]
constexpr uint32_t K_OFF = 0, V_OFF = 32768, Q_OFF = 65536, P_OFF = 81920, SMEM_TOTAL = 81920 + 65536;

// swizzled byte offset of 16-byte chunk `ch` of 128-byte row `row`
__host__ __device__ inline uint32_t sw128(uint32_t row, uint32_t ch) { return row * 128 + ((ch ^ (row & 7)) << 4); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
             const uint8_t* __restrict__ q_img, Args a, float* __restrict__ outS, float* __restrict__ outO) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_ld, bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t base = ptx::smem_u32(sm);
  if (t == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar_ld), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_s), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_o), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc_512(ptx::smem_u32(&tmem_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  cluster.sync();
  // ---- K and V rows: key = rank * 128 + t, written into CTA 0's shared memory (local for rank 0, DSMEM for rank 1)
  {
    uint8_t* dst = cluster.map_shared_rank(sm, 0);
    const int key = rank * 128 + t;
    const uint4* ks = reinterpret_cast<const uint4*>(k + static_cast<size_t>(key) * 48);
    const uint4* vs = reinterpret_cast<const uint4*>(v + static_cast<size_t>(key) * 48);
#pragma unroll
    for (int ch = 0; ch < 6; ++ch) {
      *reinterpret_cast<uint4*>(dst + K_OFF + sw128(key, ch)) = ks[ch];
      *reinterpret_cast<uint4*>(dst + V_OFF + sw128(key, ch)) = vs[ch];
    }
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  cluster.sync();
  if (rank != 0) {       // CTA 1 is done (it must not exit before CTA 0 has stopped using the cluster... it only received)
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc_512(tmem);
    return;
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  // ---- Q: tensor memory (packed) or shared memory (bulk copy of the host-made image)
  if (a.a_mode == 0) {
    uint32_t pk[24];
    const __nv_bfloat16* qr = q + static_cast<size_t>(t) * 48;
#pragma unroll
    for (int i = 0; i < 24; ++i) pk[i] = pack2(__bfloat162float(qr[2 * i]), __bfloat162float(qr[2 * i + 1]));
    ptx::tmem_st16(lane_addr + QCOL, pk);
    tmem_st8(lane_addr + QCOL + 16, pk + 16);
    ptx::tmem_st_wait();
  } else if (t == 0) {
    ptx::mbar_expect_tx(ptx::smem_u32(&bar_ld), 16384);
    bulk_load(base + Q_OFF, q_img, 16384, ptx::smem_u32(&bar_ld));
    ptx::mbar_wait(ptx::smem_u32(&bar_ld), 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (t == 0) {
    const uint32_t idesc = make_idesc(128, 256);
    const uint64_t kd = make_desc_sw128(base + K_OFF);
    for (int ks = 0; ks < 3; ++ks) {
      if (a.a_mode == 0) umma_bf16_ts(tmem, tmem + QCOL + 8 * ks, kd + 2 * ks, idesc, ks ? 1u : 0u);
      else ptx::umma_bf16(tmem, make_desc_sw128(base + Q_OFF) + 2 * ks, kd + 2 * ks, idesc, ks ? 1u : 0u);
    }
    ptx::umma_commit(ptx::smem_u32(&bar_s));
  }
  ptx::mbar_wait(ptx::smem_u32(&bar_s), 0);
  ptx::tc_fence_after();
  // ---- S out, row max
  float mx = -INFINITY;
  for (int c0 = 0; c0 < 256; c0 += 32) {
    float s[32];
    ptx::tmem_ld32(lane_addr + c0, s);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) { outS[t * 256 + c0 + i] = s[i]; mx = fmaxf(mx, s[i]); }
  }
  // ---- P = exp2(S - max) as bf16: in place over the S columns (tensor memory) or into shared memory
  for (int c0 = 0; c0 < 256; c0 += 32) {
    float s[32];
    ptx::tmem_ld32(lane_addr + c0, s);
    ptx::tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack2(exp2f(s[2 * i] - mx), exp2f(s[2 * i + 1] - mx));
    if (a.a_mode == 0) {
      ptx::tmem_st16(lane_addr + c0 / 2, pk);
    } else {
      const int kblk = c0 / 64, ch0 = (c0 % 64) / 8;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(sm + P_OFF + kblk * 16384 + sw128(t, ch0 + j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
  }
  if (a.a_mode == 0) ptx::tmem_st_wait(); else ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (t == 0) {
    const uint32_t idesc = make_idesc(128, 48) | (a.vmajor ? (1u << 16) : 0u);
    for (int ks = 0; ks < 16; ++ks) {
      const uint64_t vd = make_desc(base + V_OFF + ks * 2048, a.lbo, a.sbo);
      if (a.a_mode == 0) umma_bf16_ts(tmem + OCOL, tmem + 8 * ks, vd, idesc, ks ? 1u : 0u);
      else ptx::umma_bf16(tmem + OCOL, make_desc_sw128(base + P_OFF + (ks / 4) * 16384) + 2 * (ks % 4), vd, idesc, ks ? 1u : 0u);
    }
    ptx::umma_commit(ptx::smem_u32(&bar_o));
  }
  ptx::mbar_wait(ptx::smem_u32(&bar_o), 0);
  ptx::tc_fence_after();
  {
    float o[32], o2[16];
    ptx::tmem_ld32(lane_addr + OCOL, o);
    ptx::tmem_ld16(lane_addr + OCOL + 32, o2);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) outO[t * 48 + i] = o[i];
    for (int i = 0; i < 16; ++i) outO[t * 48 + 32 + i] = o2[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc_512(tmem);
}

int main() {
  const int NQ = 128, NK = 256, D = 48;
  std::vector<__nv_bfloat16> hq(NQ * D), hk(NK * D), hv(NK * D);
  std::vector<float> fq(NQ * D), fk(NK * D), fv(NK * D);
  srand(11);
  auto rnd = [] { return (rand() % 2001 - 1000) / 1000.f; };
  for (int i = 0; i < NQ * D; ++i) { hq[i] = __float2bfloat16(rnd() * 0.8f); fq[i] = __bfloat162float(hq[i]); }
  for (int i = 0; i < NK * D; ++i) { hk[i] = __float2bfloat16(rnd()); fk[i] = __bfloat162float(hk[i]); }
  for (int i = 0; i < NK * D; ++i) { hv[i] = __float2bfloat16(rnd()); fv[i] = __bfloat162float(hv[i]); }
  std::vector<uint8_t> qimg(16384, 0);
  for (int r = 0; r < NQ; ++r)
    for (int ch = 0; ch < 6; ++ch) memcpy(qimg.data() + sw128(r, ch), &hq[r * D + ch * 8], 16);
  // host reference
  std::vector<float> S(NQ * NK), O(NQ * D);
  for (int i = 0; i < NQ; ++i) {
    float mx = -INFINITY;
    for (int j = 0; j < NK; ++j) {
      double s = 0;
      for (int d = 0; d < D; ++d) s += static_cast<double>(fq[i * D + d]) * fk[j * D + d];
      S[i * NK + j] = static_cast<float>(s);
      mx = fmaxf(mx, S[i * NK + j]);
    }
    for (int d = 0; d < D; ++d) {
      double o = 0;
      for (int j = 0; j < NK; ++j) o += static_cast<double>(__bfloat162float(__float2bfloat16(exp2f(S[i * NK + j] - mx)))) * fv[j * D + d];
      O[i * D + d] = static_cast<float>(o);
    }
  }
  __nv_bfloat16 *dq, *dk, *dv;
  uint8_t* dqi;
  float *dS, *dO;
  cudaMalloc(&dq, hq.size() * 2); cudaMalloc(&dk, hk.size() * 2); cudaMalloc(&dv, hv.size() * 2);
  cudaMalloc(&dqi, qimg.size()); cudaMalloc(&dS, S.size() * 4); cudaMalloc(&dO, O.size() * 4);
  cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dk, hk.data(), hk.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dv, hv.data(), hv.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dqi, qimg.data(), qimg.size(), cudaMemcpyHostToDevice);
  const size_t smem = SMEM_TOTAL + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const Args variants[] = {{0, 1024, 1024, 1}, {1, 1024, 1024, 1}, {0, 16, 1024, 1}, {0, 1024, 16, 1}, {1, 16, 1024, 1}, {1, 1024, 16, 1}};
  int ok_main = 1;
  for (const Args& a : variants) {
    cudaMemset(dS, 0, S.size() * 4); cudaMemset(dO, 0, O.size() * 4);
    probe_kernel<<<2, 128, smem>>>(dq, dk, dv, dqi, a, dS, dO);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("a_mode=%d lbo=%d sbo=%d: CUDA error %s\n", a.a_mode, a.lbo, a.sbo, cudaGetErrorString(e)); return 4; }
    std::vector<float> gS(S.size()), gO(O.size());
    cudaMemcpy(gS.data(), dS, gS.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(gO.data(), dO, gO.size() * 4, cudaMemcpyDeviceToHost);
    double eS = 0, eO = 0, eS_lo = 0, eS_hi = 0;
    for (int i = 0; i < NQ; ++i)
      for (int j = 0; j < NK; ++j) {
        const double d = fabs(gS[i * NK + j] - S[i * NK + j]);
        eS = fmax(eS, d);
        if (j < 128) eS_lo = fmax(eS_lo, d); else eS_hi = fmax(eS_hi, d);
      }
    for (size_t i = 0; i < O.size(); ++i) eO = fmax(eO, fabs(gO[i] - O[i]));
    printf("A in %s, V desc lbo=%d sbo=%d b_major=%s:  S max err %.5f (local keys %.5f, DSMEM keys %.5f)   O max err %.5f  %s\n",
           a.a_mode == 0 ? "TMEM" : "smem", a.lbo, a.sbo, a.vmajor ? "MN" : "K", eS, eS_lo, eS_hi, eO,
           (eS < 2e-3 && eO < 5e-2) ? "OK" : "MISMATCH");
    if (&a == &variants[0] && !(eS < 2e-3 && eO < 5e-2)) ok_main = 0;
  }
  return ok_main ? 0 : 1;
}
