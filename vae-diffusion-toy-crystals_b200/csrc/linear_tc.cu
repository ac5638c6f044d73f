// linear_tc.cu — nn.Linear as a tcgen05 GEMM for the latent diffusion prior.
//
// Replaces the F.linear calls of DiffusionPriorFiLM / FiLMResBlock (reference:
// src/toycrystals/models/diffusion_prior.py:43-54 fc1/fc2/cond, :86-97 y_fuse/t_mlp) for bf16 operands:
//
//   out[M, N] = act(A[M, K] · W[N, K]^T + bias[N])
//
// * A (activations) and W (the nn.Linear weight, already [out, in] = K-major) are bf16; TMA (SWIZZLE_128B, 64-element
//   = 128-byte K slabs) stages 128 x 64 A tiles and 256 x 64 W tiles in a 4-6 deep mbarrier ring;
// * one elected thread issues tcgen05.mma kind::f16 (M = 128 per CTA, N = 256, K = 16), fp32 accumulators in TMEM,
//   two 256-column sets so the epilogue of tile i overlaps the main loop of tile i+1;
// * CTA pairs (cluster of 2, cta_group::2): the pair shares the N tile, each CTA loads only half of the W tile
//   (128 rows) and the leader's MMA (M = 256) reads it half from each CTA's shared memory;
// * epilogue warps (4) read TMEM with tcgen05.ld and apply bias, SiLU (fc1) or the fp32 residual accumulate (fc2);
// * persistent grid, one CTA per SM.
#include "linear_tc.cuh"

#include <cstdlib>

#include "tc_ptx.cuh"

namespace tcs {

constexpr int LT_BM = 128, LT_BN = 256, LT_BK = 64;
constexpr int LT_THREADS = 64 + 128;   // TMA warp, MMA warp, 4 epilogue warps
constexpr int LT_MAX_STAGES = 8;

struct __align__(8) LtBarriers {
  uint64_t full[LT_MAX_STAGES];
  uint64_t empty[LT_MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float lt_silu(float y) {   // y * sigmoid(y) = h + h tanh(h), h = y / 2 (one MUFU)
  const float h = 0.5f * y;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ uint32_t lt_pack(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int CG>
__device__ __forceinline__ void lt_tile_to_mn(int tile, int n_mtiles, int& mt, int& nt) {
  if (CG == 2) {   // the two CTAs of a pair (consecutive tile indices) share the N tile
    const int pt = tile >> 1, half = n_mtiles >> 1;
    nt = pt / half;
    mt = ((pt - nt * half) << 1) | (tile & 1);
  } else {
    nt = tile / n_mtiles;
    mt = tile - nt * n_mtiles;
  }
}

template <int CG>
__global__ void __launch_bounds__(LT_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                 const LinearTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ LtBarriers bars;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int n_tiles = p.n_mtiles * p.n_ntiles;
  constexpr int NB = LT_BN / CG;                 // W rows this CTA stages per K block
  constexpr uint32_t B_BYTES = NB * LT_BK * 2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapA);
    ptx::prefetch_tmap(&mapW);
    for (int s = 0; s < p.nstage; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars.full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(ptx::smem_u32(&bars.tmem_full[a]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.tmem_empty[a]), 4 * CG);
    }
    ptx::fence_barrier_init();
  }
  const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  if (warp == 1) {
    if (CG == 2) ptx::tmem_alloc_512_2sm(ptx::smem_u32(&bars.tmem_base));
    else ptx::tmem_alloc_512(ptx::smem_u32(&bars.tmem_base));
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int mt, nt;
      lt_tile_to_mn<CG>(tile, p.n_mtiles, mt, nt);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        ptx::mbar_wait(ptx::smem_u32(&bars.empty[stage]), phase ^ 1);
        if (lane == 0) {
          const uint32_t full = ptx::smem_u32(&bars.full[stage]);
          const uint32_t a_dst = smem_base + stage * p.stage_bytes;
          const uint32_t b_dst = a_dst + p.a_bytes;
          if (CG == 2) {   // both CTAs' loads complete on the leader's barrier; the leader arms it for both
            if (cta_rank == 0) ptx::mbar_expect_tx(full, 2 * (p.a_bytes + B_BYTES));
            ptx::tma_load_2d_2sm(a_dst, &mapA, full, kb * LT_BK, mt * LT_BM);
            ptx::tma_load_2d_2sm(b_dst, &mapW, full, kb * LT_BK, nt * LT_BN + static_cast<int>(cta_rank) * NB);
          } else {
            ptx::mbar_expect_tx(full, p.a_bytes + B_BYTES);
            ptx::tma_load_2d(a_dst, &mapA, full, kb * LT_BK, mt * LT_BM);
            ptx::tma_load_2d(b_dst, &mapW, full, kb * LT_BK, nt * LT_BN);
          }
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(p.nstage)) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ============================== MMA issuer (leader CTA only when CG == 2) ====
    constexpr uint32_t idesc = make_idesc(LT_BM * CG, LT_BN);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      ptx::mbar_wait(ptx::smem_u32(&bars.tmem_empty[acc]), acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * LT_BN;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        ptx::mbar_wait(ptx::smem_u32(&bars.full[stage]), phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a_base = smem_base + stage * p.stage_bytes;
          const uint64_t adesc = make_desc_sw128(a_base), bdesc = make_desc_sw128(a_base + p.a_bytes);
#pragma unroll
          for (int k = 0; k < LT_BK / 16; ++k) {   // 32 bytes per K = 16 step inside the 128-byte swizzle row
            const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
            if (CG == 2) ptx::umma_bf16_2sm(d_tmem, adesc + k * 2, bdesc + k * 2, idesc, accum);
            else ptx::umma_bf16(d_tmem, adesc + k * 2, bdesc + k * 2, idesc, accum);
          }
          if (CG == 2) {
            ptx::umma_commit_2sm(ptx::smem_u32(&bars.empty[stage]));
            if (kb == p.kblocks - 1) ptx::umma_commit_2sm(ptx::smem_u32(&bars.tmem_full[acc]));
          } else {
            ptx::umma_commit(ptx::smem_u32(&bars.empty[stage]));
            if (kb == p.kblocks - 1) ptx::umma_commit(ptx::smem_u32(&bars.tmem_full[acc]));
          }
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(p.nstage)) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2) {
    // ============================== epilogue: warp q reads TMEM lane quarter q = warp % 4 ====
    const int q = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    const bool silu = (p.flags & LIN_SILU) != 0, f32 = (p.flags & LIN_OUT_F32) != 0, accum = (p.flags & LIN_ACCUM) != 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int mt, nt;
      lt_tile_to_mn<CG>(tile, p.n_mtiles, mt, nt);
      const int row = mt * LT_BM + q * 32 + lane;
      const bool live = row < p.M;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * LT_BN;
      ptx::mbar_wait(ptx::smem_u32(&bars.tmem_full[acc]), acc_phase);
      ptx::tc_fence_after();
      float vbuf[2][32];
      ptx::tmem_ld32(tbase, vbuf[0]);
#pragma unroll
      for (int c = 0; c < LT_BN / 32; ++c) {
        ptx::tmem_ld_wait();
        if (c + 1 < LT_BN / 32) ptx::tmem_ld32(tbase + (c + 1) * 32, vbuf[(c + 1) & 1]);
        float* v = vbuf[c & 1];
        const int col = nt * LT_BN + c * 32;
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
        }
        if (silu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = lt_silu(v[i]);
        }
        if (live && f32) {
          float* o = static_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 r = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if (accum) {
              const float4 a4 = *reinterpret_cast<const float4*>(o + i);
              r.x += a4.x; r.y += a4.y; r.z += a4.z; r.w += a4.w;
            }
            *reinterpret_cast<float4*>(o + i) = r;
          }
        } else if (live) {
          uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col);
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            o[i / 8] = make_uint4(lt_pack(v[i], v[i + 1]), lt_pack(v[i + 2], v[i + 3]), lt_pack(v[i + 4], v[i + 5]),
                                  lt_pack(v[i + 6], v[i + 7]));
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) ptx::mbar_arrive_rank0(ptx::smem_u32(&bars.tmem_empty[acc]));
        else ptx::mbar_arrive(ptx::smem_u32(&bars.tmem_empty[acc]));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync_all();   // the leader's MMAs read the peer's shared memory / write its TMEM
  ptx::tc_fence_after();
  if (warp == 1) {
    if (CG == 2) ptx::tmem_dealloc_512_2sm(tmem_base);
    else ptx::tmem_dealloc_512(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled lt_get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static int lt_encode_2d(CUtensorMap* map, const __nv_bfloat16* base, int rows, int cols, int pitch, int box_rows) {
  PFN_encodeTiled encode = lt_get_encode();
  if (!encode) return fail(TCS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch) * 2};
  cuuint32_t box[2] = {LT_BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TCS_ERR_CUDA, "cuTensorMapEncodeTiled(linear) failed: " + std::to_string(r));
  return TCS_OK;
}

int linear_tc_make_plan(LinearTcPlan* plan, const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N,
                        int K, const float* bias, void* out, int ldo, int flags, int sm_count) {
  if (M < 1 || N % LT_BN || K % LT_BK || K < LT_BK)
    return fail(TCS_ERR_UNSUPPORTED, "linear_tc: needs N % 256 == 0 and K % 64 == 0 (got N=" + std::to_string(N) +
                                         ", K=" + std::to_string(K) + ")");
  if ((lda % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return fail(TCS_ERR_BAD_ARGUMENT, "linear_tc: operands must be 16-byte aligned with 16-byte row pitches");
  if ((flags & LIN_ACCUM) && !(flags & LIN_OUT_F32)) return fail(TCS_ERR_BAD_ARGUMENT, "linear_tc: accumulate needs fp32 output");
  LinearTcPlan& pl = *plan;
  pl = LinearTcPlan();
  {
    const char* e = getenv("TCS_CG");   // 1 = single-CTA MMA (A/B switch)
    pl.cg = (e && atoi(e) == 1) ? 1 : 2;
  }
  LinearTcParams& p = pl.p;
  p.M = M; p.N = N; p.K = K;
  p.n_mtiles = (M + LT_BM - 1) / LT_BM;
  if (pl.cg == 2 && (p.n_mtiles & 1)) ++p.n_mtiles;   // whole CTA pairs: the extra tile is all out-of-bounds rows
  p.n_ntiles = N / LT_BN;
  p.kblocks = K / LT_BK;
  p.a_bytes = LT_BM * LT_BK * 2;
  p.stage_bytes = p.a_bytes + (LT_BN / pl.cg) * LT_BK * 2;
  const size_t budget = 227 * 1024 - 2048 - 1024;
  p.nstage = static_cast<int>(budget / p.stage_bytes);
  if (p.nstage > LT_MAX_STAGES) p.nstage = LT_MAX_STAGES;
  pl.smem = static_cast<size_t>(p.nstage) * p.stage_bytes + 1024;
  p.bias = bias; p.out = out; p.ldo = ldo; p.flags = flags;
  const int tiles = p.n_mtiles * p.n_ntiles;
  pl.grid = tiles < sm_count ? tiles : sm_count;
  if (pl.cg == 2) pl.grid &= ~1;
  TCS_CHECK(lt_encode_2d(&pl.mapA, A, M, K, lda, LT_BM));
  TCS_CHECK(lt_encode_2d(&pl.mapW, W, N, K, ldw, LT_BN / pl.cg));
  pl.valid = true;
  return TCS_OK;
}

template <int CG>
static int lt_launch_t(const LinearTcPlan& pl, cudaStream_t st) {
  auto kern = linear_tc_kernel<CG>;
  static bool attr_done = false;
  if (!attr_done) {
    TCS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pl.grid); cfg.blockDim = dim3(LT_THREADS); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  int na = 0;
  if (CG == 2) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  TCS_CUDA(cudaLaunchKernelEx(&cfg, kern, pl.mapA, pl.mapW, pl.p));
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

int linear_tc_launch(const LinearTcPlan& pl, cudaStream_t st) {
  if (!pl.valid) return fail(TCS_ERR_STATE, "linear_tc_launch: plan not built");
  return pl.cg == 2 ? lt_launch_t<2>(pl, st) : lt_launch_t<1>(pl, st);
}

}  // namespace tcs
