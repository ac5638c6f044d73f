/*
 * tcs_prior.h — C ABI of the latent-diffusion-prior sampler and the CondVAE decoder in libtcs.so
 * (BASELINE configs[3]: prior MLP sampler T=1000, width 1024, z_dim 32, beta_end 0.05, then
 * conditional VAE decode to 64x64).
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *
 *   tcs_prior_create / _set_weight / _finalize_weights / _destroy
 *        <- DiffusionPriorFiLM.__init__ + load_state_dict   src/toycrystals/models/diffusion_prior.py:62-106
 *           (built with n_blocks=8, y_cat_emb_dim=64 at       scripts/train_diffusion_prior.py:196-204)
 *   tcs_prior_eps
 *        <- DiffusionPriorFiLM.forward                        diffusion_prior.py:108-127
 *   tcs_prior_ddim_sample
 *        <- DiffusionSchedule.linear + .ddim_sample (eta=0)   diffusion_prior.py:177-188, 200-252
 *   tcs_prior_schedule_host / tcs_prior_timesteps_host
 *        <- alpha_bars (:179-181) and the DDIM timestep subset (:217-222), exactly as the kernels use them
 *   tcs_vae_create / _set_weight / _finalize_weights / _destroy / tcs_vae_decode
 *        <- CondVAE.__init__ / load_state_dict / decode (eval)  src/toycrystals/models/vae.py:9-43, 62-70
 *           and the un-standardisation z = z_norm * z_std + z_mean of save_diffusion_samples
 *                                                             scripts/train_diffusion_prior.py:88-95
 *
 * Conventions as in tcs.h: plain C types, DEVICE pointers unless the name ends in _host, caller-owned buffers,
 * asynchronous on `stream`, negative tcs_status on error with the text in tcs_last_error().  No CPU fallback.
 */
#ifndef TCS_PRIOR_H_
#define TCS_PRIOR_H_

#include "tcs.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tcs_prior tcs_prior;
typedef struct tcs_vae tcs_vae;

typedef struct tcs_prior_config {
  int32_t z_dim;          /* <= 32, multiple of 4                                                        */
  int32_t n_types;
  int32_t y_cont_dim;     /* <= 16                                                                        */
  int32_t t_emb_dim;      /* multiple of 16 (64 in the reference CLI)                                     */
  int32_t width;          /* 256, 512, 1024 or 2048 (README figure: 1024)                                 */
  int32_t n_blocks;       /* FiLM residual blocks (8 at the reference call site)                          */
  int32_t y_cat_emb_dim;  /* 32, 64 or 128 (64 at the reference call site)                                */
  int32_t T;              /* DiffusionSchedule.linear(T, beta_start, beta_end)                            */
  double beta_start;
  double beta_end;
  int32_t precision;      /* tcs_precision: TCS_BF16 = tcgen05 GEMMs (bf16 operands, fp32 accumulate and   */
                          /* fp32 residual stream), TCS_FP32 = FFMA GEMMs (the 1e-4 parity mode)           */
  int32_t device;
  int32_t use_graph;      /* 1 = one DDIM step is captured in a CUDA graph and replayed                    */
} tcs_prior_config;

/* README "settings used for the figures": z 32, 4 types, 4 cont, t_emb 64, width 1024, 8 blocks, emb 64,
 * T 1000, beta 1e-4 .. 0.05, bf16, device 0, graph on. */
void tcs_prior_default_config(tcs_prior_config* cfg);

int tcs_prior_create(tcs_prior** out, const tcs_prior_config* cfg);
void tcs_prior_destroy(tcs_prior* h);
/* key spelled as in DiffusionPriorFiLM.state_dict(); fp32, host or device memory, copied before returning */
int tcs_prior_set_weight(tcs_prior* h, const char* key, const float* data, const int64_t* shape, int32_t ndim);
int tcs_prior_finalize_weights(tcs_prior* h);

/* eps_out[n, z_dim] = model(z_t[n, z_dim], t[n] (int64 timesteps), y_cat[n] int64, y_cont[n, y_cont_dim]) */
int tcs_prior_eps(tcs_prior* h, const float* z_t, const int64_t* t, const int64_t* y_cat, const float* y_cont, int32_t n,
                  float* eps_out, void* stream);

typedef struct tcs_ddim_args {
  int32_t n;                    /* samples of this call                                                   */
  int32_t n_steps;              /* requested steps; the schedule keeps unique_consecutive(round(linspace)) */
  const int64_t* y_cat;         /* [n]                                                                     */
  const float* y_cont;          /* [n, y_cont_dim]                                                         */
  const float* z_init;          /* [n, z_dim] or NULL -> Philox N(0,1) keyed (seed, global index)          */
  uint64_t seed;
  uint64_t global_index_offset; /* index of sample 0 of this call in the whole job (sharding)              */
  float* z0_out;                /* [n, z_dim] normalised latents                                           */
  float* trace_eps;             /* NULL or [steps_run, n, z_dim]: eps of every evaluation                  */
  float* trace_z;               /* NULL or [steps_run, n, z_dim]: the z_t every evaluation saw             */
} tcs_ddim_args;

int tcs_prior_ddim_sample(tcs_prior* h, const tcs_ddim_args* args, void* stream);

/* Host-side schedule exactly as the kernels use it (tests compare them with torch):
 * alpha_bars_host[T] = cumprod(1 - linspace(beta_start, beta_end, T)) in fp32 */
int tcs_prior_schedule_host(int32_t T, double beta_start, double beta_end, float* alpha_bars_host);
/* ts_host[<= n_steps] = unique_consecutive(round(linspace(T-1, 0, n_steps))); *count_host = entries written */
int tcs_prior_timesteps_host(int32_t T, int32_t n_steps, int64_t* ts_host, int32_t* count_host);
int64_t tcs_prior_launch_count(const tcs_prior* h);

/* ---- CondVAE decoder ------------------------------------------------------------------------------------ */
typedef struct tcs_vae_config {
  int32_t z_dim;       /* <= 32 */
  int32_t n_types;
  int32_t y_cont_dim;
  int32_t device;
  int32_t precision;   /* tcs_precision: TCS_BF16 = the three wide ConvTranspose2d stages on tcgen05 (bf16 activations
                          and weights, fp32 accumulate); TCS_FP32 = FFMA kernels throughout (the parity mode) */
} tcs_vae_config;

int tcs_vae_create(tcs_vae** out, const tcs_vae_config* cfg);
void tcs_vae_destroy(tcs_vae* h);
/* key spelled as in CondVAE.state_dict(); encoder tensors (enc.*, enc_fc.*, mu.*, logvar.*) are accepted and ignored */
int tcs_vae_set_weight(tcs_vae* h, const char* key, const float* data, const int64_t* shape, int32_t ndim);
int tcs_vae_finalize_weights(tcs_vae* h);
/* x_out[n,1,64,64] = decode(z, y_cat, y_cont) in eval mode; when z_mean / z_std ([z_dim], both or neither) are given
 * the input is a normalised latent and z * z_std + z_mean is applied first */
int tcs_vae_decode(tcs_vae* h, const float* z, const int64_t* y_cat, const float* y_cont, int32_t n, const float* z_mean,
                   const float* z_std, float* x_out, void* stream);
int64_t tcs_vae_launch_count(const tcs_vae* h);

/* Timing hook for bench.py: runs, `reps` times each and back to back on the library's stream with CUDA events around
 * the group, the four per-step kernels of block 0 on the first n rows of the workspace left by the last
 * tcs_prior_ddim_sample call (which must have covered >= n rows): ms_host[0] = fc1 GEMM (tcgen05 in bf16 mode),
 * [1] = fc2 GEMM, [2] = LayerNorm+FiLM, [3] = tail (out_norm/out_proj/DDIM/in_proj).  Average ms per launch.
 * The residual stream is scratch afterwards.  Synchronous. */
int tcs_prior_profile(tcs_prior* h, int32_t n, int32_t reps, float* ms_host);

/* ---- test hook (tests/ only) ------------------------------------------------------------------------------ */
/* One dense layer in isolation, synchronous: out[M,N] = act(A[M,K] W[N,K]^T + bias) (+ out when accumulate).
 * A, W, bias (nullable), out are fp32 device tensors; engine TCS_ENGINE_TCGEN05 rounds A and W to bf16 and runs the
 * tensor-core GEMM (N % 256 == 0, K % 64 == 0), TCS_ENGINE_SIMT runs the FFMA GEMM (N % 32 == 0, K % 16 == 0).
 * bf16_out = 1 stores the result as bf16 (as between fc1 and fc2) before it is widened back into `out`. */
int tcs_debug_linear(int32_t engine, int32_t M, int32_t N, int32_t K, const float* A, const float* W, const float* bias,
                     float* out, int32_t silu, int32_t accumulate, int32_t bf16_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TCS_PRIOR_H_ */
