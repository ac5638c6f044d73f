// prior.cuh — host launchers of the CUDA-core kernels of the latent diffusion prior and the CondVAE decoder
// (kernels_prior.cu).  The tensor-core GEMM is linear_tc.cuh.
#pragma once
#include "common.cuh"
#include "linear_tc.cuh"

namespace tcs {

constexpr float LN_EPS = 1e-5f;      // nn.LayerNorm default (diffusion_prior.py:42,105)
constexpr int PRIOR_MAX_WIDTH = 2048;
constexpr int PRIOR_MAX_Z = 32;

// ---- fp32 GEMM on CUDA cores: out[M,N] = act(A[M,K] W[N,K]^T + bias); flags = LinFlags ------------------------
// A fp32 (row pitch lda), W fp32 (row pitch ldw, the nn.Linear layout).  N % 64 == 0, K % 16 == 0.
int launch_linear_simt(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const float* bias, void* out,
                       int ldo, int flags, cudaStream_t st);

// ---- condition / time features (diffusion_prior.py:11-25, 76-97, 110-116) ---------------------------------------
struct PriorEmbedWeights {
  const float* cat_emb;          // y_cat_emb.weight [n_types, E]
  const float *cm0_w, *cm0_b;    // y_cont_mlp.0 [E, ycd]
  const float *cm2_w, *cm2_b;    // y_cont_mlp.2 [E, E]
  int n_types, y_cont_dim, E;
};
// ycat[n][2E] = [y_cat_emb[y_cat], y_cont_mlp(y_cont)]
int launch_prior_ycat(const PriorEmbedWeights& w, const int64_t* y_cat, const float* y_cont, int n, float* out, cudaStream_t st);
// te[r][dim] = [sin(t_r f_i), cos(t_r f_i)] (+ a zero column when dim is odd); t int64 or (t_is_i32) int32; freqs [dim/2]
int launch_prior_time_features(const void* t, int t_is_i32, const float* freqs, int rows, int dim, float* te, cudaStream_t st);

// ---- LayerNorm + FiLM: u = LN(h) (1 + gamma) + beta  (FiLMResBlock.forward :49-52) ------------------------------
// gamma = film_row[r][off + c] (+ film_step[step][off + c]),  beta = the same at off + W + c,  off = blk * 2W.
// film_row (stored like the output: bf16 in bf16 mode, it is the dominant HBM stream of the kernel) may be null (all-step
// tables only) and film_step (fp32) may be null (per-row FiLM only).
template <typename TO>
int launch_ln_film(const float* h, int n, int W, const float* ln_w, const float* ln_b, const TO* film_row, int film_ld,
                   const float* film_step, int film_step_ld, const int* step_ptr, int off, TO* out, cudaStream_t st);

// ---- end of a network evaluation: out_norm + out_proj (:125-126), the DDIM update (:228-250) and in_proj of the next
// evaluation (:118), one kernel --------------------------------------------------------------------------------------
struct DdimCoef {      // one row per DDIM step, host-computed in the reference's fp32 operation order
  float s1m_t;         // sqrt(1 - abar_t)
  float sa_t_eps;      // sqrt(abar_t) + 1e-8
  float sa_prev;       // sqrt(abar_prev)          (unused on the last step)
  float s1m_prev;      // sqrt(1 - abar_prev)
  int last;            // 1 = this is the last step: z <- z0_pred
  int t;               // integer timestep (for reference / debugging)
  int pad0, pad1;
};
enum PriorTailMode : int { TAIL_INIT = 0, TAIL_EPS = 1, TAIL_DDIM = 2 };
struct PriorTailArgs {
  int mode;                 // INIT: h = in_proj(z) only; EPS: eps_out = out_proj(LN(h)) only; DDIM: eps, update z, h = in_proj(z)
  int n, W, zd;
  float* h;                 // [n, W] residual stream (read for eps, rewritten by in_proj)
  float* z;                 // [n, zd] latent state (fp32)
  const float *on_w, *on_b; // out_norm
  const float *op_w, *op_b; // out_proj [zd, W]
  const float *ip_w, *ip_b; // in_proj  [W, zd]
  const DdimCoef* coef;     // device table
  const int* step_ptr;      // device step counter
  float* eps_out;           // EPS: [n, zd]
  float* z_out;             // DDIM: receives z0 on the last step (may alias z)
  float* trace_eps;         // DDIM: null or [S, trace_n, zd]
  float* trace_z;           // DDIM: null or [S, trace_n, zd] = the z every evaluation saw
  int trace_n;              // rows per step in the traces (>= n when the job runs in chunks)
};
int launch_prior_tail(const PriorTailArgs& a, cudaStream_t st);

// z[i][:] ~ N(0,1) keyed (seed, gidx0 + i, word 0)
int launch_prior_philox(float* z, int n, int zd, unsigned long long seed, unsigned long long gidx0, cudaStream_t st);
// fp32 -> bf16 copy (row pitch aware)
int launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, size_t count, cudaStream_t st);

// ---- CondVAE decoder (vae.py:36-43, 62-70) ----------------------------------------------------------------------
// h0 NHWC [n,4,4,256] = dec_fc([z', onehot(y_cat), y_cont]),  z' = z * z_std + z_mean when z_mean/z_std are given
int launch_vae_dec_fc(const float* z, const int64_t* y_cat, const float* y_cont, const float* z_mean, const float* z_std,
                      const float* w /*[4096, zd+n_types+ycd]*/, const float* b, int n, int zd, int n_types, int ycd,
                      void* h0 /*fp32 or bf16*/, int out_bf16, cudaStream_t st);
// ConvTranspose2d(k=4, s=2, p=1) + ReLU as four parity-class GEMMs: in NHWC [n,Hi,Hi,Ci] -> out NHWC [n,2Hi,2Hi,Co];
// wpacked fp32 [4 parity][Co][4 taps * Ci]
int launch_vae_convt(const float* in, const float* wpacked, const float* bias, int n, int Hi, int Ci, int Co, float* out,
                     cudaStream_t st);
void vae_convt_pack_weights(const float* w /*[Ci,Co,4,4]*/, int Ci, int Co, float* out_host);
// last layer: ConvTranspose2d(32 -> 1) + Sigmoid: in NHWC [n,32,32,32] -> x [n,64,64]; wpacked [4 parity][4 taps][32]
int launch_vae_convt_out(const void* in /*fp32 or bf16*/, int in_bf16, const float* wpacked, float bias, int n, float* x,
                         cudaStream_t st);

}  // namespace tcs
