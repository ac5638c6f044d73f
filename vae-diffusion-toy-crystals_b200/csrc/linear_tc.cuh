// linear_tc.cuh — host interface of the tcgen05 GEMM used by the latent diffusion prior (linear_tc.cu):
//   out[M,N] = act(A[M,K] · W[N,K]^T + bias)     A, W bf16 K-major (W is the nn.Linear weight layout), fp32 accumulate.
#pragma once
#include "common.cuh"

namespace tcs {

enum LinFlags : int {
  LIN_SILU = 1,        // SiLU on (acc + bias)
  LIN_OUT_F32 = 2,     // fp32 output (default bf16)
  LIN_ACCUM = 4        // fp32 output only: out += acc + bias (the residual stream h of the FiLM blocks)
};

struct LinearTcParams {
  int M, N, K;
  int n_mtiles, n_ntiles, kblocks;   // 128-row M tiles (padded to an even count for CTA pairs), 256-col N tiles, K/64
  int nstage;
  uint32_t a_bytes, stage_bytes;
  const float* bias;                 // [N] or null
  void* out;
  int ldo;                           // elements between output rows
  int flags;
};

struct LinearTcPlan {
  CUtensorMap mapA, mapW;
  LinearTcParams p;
  int cg;        // 1 = one CTA per MMA, 2 = CTA pairs (cta_group::2, W tile split across the pair)
  int grid;
  size_t smem;
  bool valid = false;
};

// A: bf16 [M, K] with row pitch lda elements; W: bf16 [N, K] with row pitch ldw elements (both 16-byte aligned pitches).
// Needs N % 256 == 0 and K % 64 == 0; M is arbitrary (TMA zero-fills, the epilogue masks).
int linear_tc_make_plan(LinearTcPlan* plan, const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N,
                        int K, const float* bias, void* out, int ldo, int flags, int sm_count);
int linear_tc_launch(const LinearTcPlan& plan, cudaStream_t st);

}  // namespace tcs
