// umma_shift_test.cu — hardware probe (not part of libtcs): can ONE SWIZZLE_128B shared-memory window serve every tap
// of a 3x3 convolution through UMMA descriptor start-address shifts?
//
// The window is an (R rows x P pixels x 64 channels) TMA box (128-byte pixel rows, dense, pitch P*128 B which is NOT a
// multiple of the 1024-byte swizzle period).  An M = 128 tile is 16 image rows x 8 pixels: 8-row core groups = 8
// consecutive pixels of one image row, SBO = P*128 B (one window row), tap (ky, kx) = start + (ky*P + kx)*128 B.
// The probe runs the MMA for every tap / sub-tile and two ways of filling the descriptor's base-offset field and
// prints the max error against a host reference.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o build/umma_shift_test tools/umma_shift_test.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../vae-diffusion-toy-crystals_b200/csrc/tc_ptx.cuh"

using namespace tcs;

struct Args {
  int P, R;          // window pixels per row, rows
  int shift_px;      // start shift in pixels (ky*P + kx + sub*8)
  int mode;          // 0: base_offset = 0, 1: base_offset = (start >> 7) & 7
  int N;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Args a, float* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = a.R * a.P * 128;
  const uint32_t b_base = base + ((a_bytes + 1023u) & ~1023u);
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar_full), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_done), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc_512(ptx::smem_u32(&tmem_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(ptx::smem_u32(&bar_full), a_bytes + a.N * 128);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(base), "l"(reinterpret_cast<uint64_t>(&mapA)), "r"(ptx::smem_u32(&bar_full)), "r"(0), "r"(0), "r"(0) : "memory");
    ptx::tma_load_2d(b_base, &mapB, ptx::smem_u32(&bar_full), 0, 0);
    ptx::mbar_wait(ptx::smem_u32(&bar_full), 0);
    ptx::tc_fence_after();
    const uint32_t start = base + a.shift_px * 128;
    uint64_t ad = 0;
    ad |= static_cast<uint64_t>((start & 0x3FFFF) >> 4);
    ad |= static_cast<uint64_t>(1) << 16;
    ad |= static_cast<uint64_t>((a.P * 128) >> 4) << 32;       // SBO = one window row
    ad |= static_cast<uint64_t>(1) << 46;
    if (a.mode == 1) ad |= static_cast<uint64_t>((start >> 7) & 7) << 49;
    ad |= static_cast<uint64_t>(2) << 61;
    const uint64_t bd = make_desc_sw128(b_base);
    const uint32_t idesc = make_idesc(128, a.N);
    for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
    ptx::umma_commit(ptx::smem_u32(&bar_done));
  }
  __syncthreads();
  ptx::mbar_wait(ptx::smem_u32(&bar_done), 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < a.N; c0 += 32) {
    float v[32];
    ptx::tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * a.N + c0 + i] = v[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc_512(tmem);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess) return 2;
  PFN_encodeTiled encode = reinterpret_cast<PFN_encodeTiled>(fp);
  const int N = 96;
  int worst_fail = 0;
  for (int P : {18, 10, 34}) {
    const int R = 18;
    std::vector<__nv_bfloat16> hA(static_cast<size_t>(R) * P * 64), hB(N * 64);
    std::vector<float> fA(hA.size()), fB(hB.size());
    srand(7 + P);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dO;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap mapA, mapB;
    {
      cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(P), static_cast<cuuint64_t>(R)};
      cuuint64_t strides[2] = {128, static_cast<cuuint64_t>(128 * P)};
      cuuint32_t box[3] = {64, static_cast<cuuint32_t>(P), static_cast<cuuint32_t>(R)};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", r); return 3; }
    }
    {
      cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(N)};
      cuuint64_t strides[1] = {128};
      cuuint32_t box[2] = {64, static_cast<cuuint32_t>(N)};
      cuuint32_t es[2] = {1, 1};
      CUresult r = encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", r); return 3; }
    }
    const size_t smem = 1024 + ((R * P * 128 + 1023) & ~1023) + N * 128 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    std::vector<float> hO(128 * N);
    for (int mode = 0; mode < 2; ++mode) {
      double worst = 0;
      int nbad = 0;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx)
          for (int sub = 0; sub < ((P >= 18) ? 2 : 1); ++sub) {
            Args a{P, R, ky * P + kx + sub * 8, mode, N};
            probe_kernel<<<1, 128, smem>>>(mapA, mapB, a, dO);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("P=%d mode=%d ky=%d kx=%d sub=%d: CUDA error %s\n", P, mode, ky, kx, sub, cudaGetErrorString(e)); return 4; }
            cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
            double err = 0;
            for (int i = 0; i < 128; ++i) {
              const int w = (ky + i / 8) * P + kx + sub * 8 + (i % 8);
              for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int c = 0; c < 64; ++c) s += static_cast<double>(fA[static_cast<size_t>(w) * 64 + c]) * fB[n * 64 + c];
                err = fmax(err, fabs(s - hO[i * N + n]));
              }
            }
            if (err > 1e-3) { ++nbad; printf("  P=%d mode=%d ky=%d kx=%d sub=%d: max err %.4f\n", P, mode, ky, kx, sub, err); }
            worst = fmax(worst, err);
          }
      printf("P=%d (pitch %d B) base_offset mode %d: worst max-abs error over all taps %.6f, failing taps %d\n", P, P * 128, mode, worst, nbad);
      if (mode == 1 && nbad) worst_fail = 1;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
  }
  return worst_fail;
}
