// prior_api.cu — C ABI of the latent diffusion prior sampler and the CondVAE decoder (include/tcs_prior.h).
// BASELINE configs[3] / SURVEY 8(f) row 1.  No PyTorch types, no CPU fallback.
//
// Data flow of one DDIM run (tcs_prior_ddim_sample), n samples, S steps, width W, B blocks:
//   once:   t-chain   te[S,temb] -> t_mlp -> t_feat[S,W] -> tcond[S, B*2W] = t_feat . cond_w[:, :W]^T + cond_b   (FFMA, tiny)
//           y-chain   [emb(y_cat), y_cont_mlp(y_cont)] -> y_fuse -> y_feat[n,W] -> film[n, B*2W] = y_feat . cond_w[:, W:]^T
//                     (the condition half of every FiLM `cond` Linear is step-invariant: one tcgen05 GEMM per call)
//           h = in_proj(z_T)
//   step:   for each block:  u = LN(h)(1 + gamma) + beta  with (gamma, beta) = film[row] + tcond[step]   (ln_film_kernel)
//                            a = SiLU(u W1^T + b1)   (linear_tc, bf16 out)
//                            h += a W2^T + b2        (linear_tc, fp32 accumulate into the residual stream)
//           eps = out_proj(LN(h)); z <- DDIM(z, eps); h = in_proj(z)                                     (prior_tail_kernel)
//   The step is captured once in a CUDA graph (device-side step counter) and replayed S times.
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/tcs_prior.h"
#include "kernels.cuh"
#include "prior.cuh"

namespace tcs {

// torch.linspace(start, end, steps) in fp32 (the CUDA / scalar formula: start + step*i below the midpoint,
// end - step*(steps-1-i) above it)
static void linspace_f32(float start, float end, int steps, std::vector<float>* out) {
  out->resize(steps);
  if (steps == 1) { (*out)[0] = start; return; }
  const float step = (end - start) / static_cast<float>(steps - 1);
  const int halfway = steps / 2;
  for (int i = 0; i < steps; ++i)
    (*out)[i] = i < halfway ? start + step * static_cast<float>(i) : end - step * static_cast<float>(steps - i - 1);
}

// DiffusionSchedule.linear (:177-188): betas fp32 linspace, alphas = 1 - betas, alpha_bars = cumprod (fp32; torch's
// CPU cumprod accumulates in double and rounds every element to fp32)
static void alpha_bars_host(int T, double beta_start, double beta_end, std::vector<float>* abar) {
  std::vector<float> betas;
  linspace_f32(static_cast<float>(beta_start), static_cast<float>(beta_end), T, &betas);
  abar->resize(T);
  double acc = 1.0;
  for (int i = 0; i < T; ++i) {
    const float alpha = 1.0f - betas[i];
    acc *= static_cast<double>(alpha);
    (*abar)[i] = static_cast<float>(acc);
  }
}

// ddim_sample (:217-222): ts = unique_consecutive(round(linspace(T-1, 0, n_steps))) (round half to even)
static void ddim_timesteps_host(int T, int n_steps, std::vector<int64_t>* ts) {
  std::vector<float> lin;
  linspace_f32(static_cast<float>(T - 1), 0.0f, n_steps, &lin);
  ts->clear();
  for (int i = 0; i < n_steps; ++i) {
    const int64_t v = static_cast<int64_t>(nearbyintf(lin[i]));
    if (ts->empty() || ts->back() != v) ts->push_back(v);
  }
}

constexpr int PRIOR_CHUNK = 16384;   // samples per DDIM run (bounds the film table: 64 KiB per sample at W = 1024)
constexpr int VAE_CHUNK = 4096;      // images per decoder pass (240 KiB of fp32 activations per image)

}  // namespace tcs

using namespace tcs;

struct tcs_prior {
  tcs_prior_config cfg;
  int sm_count = 148;
  bool bf16 = true;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_pinned = nullptr;
  int64_t launches = 0;

  std::map<std::string, HostTensor> host_w;
  bool finalized = false;
  DevBuf arena, arena16;
  std::map<std::string, const float*> dw;
  std::map<std::string, const __nv_bfloat16*> dw16;
  const float *cond_w = nullptr, *cond_b = nullptr;          // all blocks: [B*2W, 2W], [B*2W]
  const __nv_bfloat16* cond_w16 = nullptr;
  const float* freqs = nullptr;
  PriorEmbedWeights ew{};

  // workspace
  int cap_n = 0, cap_s = 0;
  int64_t ws_gen = 0;
  DevBuf h, u, a, z, film, ycat, y1, yfeat, condin, te, t1, tfeat, tcond, coef, ts32, step_ctr;
  void* pinned = nullptr; size_t pinned_bytes = 0;

  cudaGraphExec_t gexec = nullptr;
  struct GKey { int n = -1; int64_t gen; const void *z0, *teps, *tz; long long tn, trow0; } gkey;
  int64_t graph_kernels = 0;

  ~tcs_prior() {
    if (gexec) cudaGraphExecDestroy(gexec);
    if (pinned) cudaFreeHost(pinned);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_out) cudaEventDestroy(ev_out);
    if (ev_pinned) cudaEventDestroy(ev_pinned);
    if (stream) cudaStreamDestroy(stream);
  }
};

struct tcs_vae {
  tcs_vae_config cfg;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  int64_t launches = 0;
  std::map<std::string, HostTensor> host_w;
  bool finalized = false;
  DevBuf arena, arena16;
  bool bf16 = false;
  int sm_count = 148;
  const __nv_bfloat16* ct_w16[3] = {};
  const float *fc_w = nullptr, *fc_b = nullptr;
  const float* ct_w[4] = {};
  const float* ct_b[3] = {};
  float out_bias = 0.f;
  DevBuf a0, a1, a2, a3;
  ~tcs_vae() {
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_out) cudaEventDestroy(ev_out);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace tcs {

template <typename H>
static int enter_h(H* h, cudaStream_t user) {
  TCS_CUDA(cudaSetDevice(h->cfg.device));
  TCS_CUDA(cudaEventRecord(h->ev_in, user));
  TCS_CUDA(cudaStreamWaitEvent(h->stream, h->ev_in, 0));
  return TCS_OK;
}
template <typename H>
static int leave_h(H* h, cudaStream_t user) {
  TCS_CUDA(cudaEventRecord(h->ev_out, h->stream));
  TCS_CUDA(cudaStreamWaitEvent(user, h->ev_out, 0));
  return TCS_OK;
}
static int check_device(int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(TCS_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libtcs has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(TCS_ERR_BAD_ARGUMENT, "bad device ordinal");
  TCS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TCS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(TCS_ERR_UNSUPPORTED, std::string("libtcs is built for sm_100a only; device is ") + prop.name);
  return TCS_OK;
}
static int store_weight(std::map<std::string, HostTensor>* m, const char* key, const float* data, const int64_t* shape,
                        int32_t ndim) {
  HostTensor t;
  size_t cnt = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); cnt *= static_cast<size_t>(shape[i]); }
  t.v.resize(cnt);
  TCS_CUDA(cudaMemcpy(t.v.data(), data, cnt * 4, cudaMemcpyDefault));
  (*m)[key] = std::move(t);
  return TCS_OK;
}

using Expected = std::vector<std::pair<std::string, std::vector<int64_t>>>;
static Expected prior_expected(const tcs_prior_config& c) {
  Expected v;
  const int64_t W = c.width, E = c.y_cat_emb_dim;
  auto lin = [&](const std::string& k, int64_t i, int64_t o) { v.push_back({k + ".weight", {o, i}}); v.push_back({k + ".bias", {o}}); };
  auto ln = [&](const std::string& k) { v.push_back({k + ".weight", {W}}); v.push_back({k + ".bias", {W}}); };
  // the `cond` Linears first and in block order: together they form one [B*2W, 2W] matrix in the arena
  for (int i = 0; i < c.n_blocks; ++i) v.push_back({"blocks." + std::to_string(i) + ".cond.weight", {2 * W, 2 * W}});
  for (int i = 0; i < c.n_blocks; ++i) v.push_back({"blocks." + std::to_string(i) + ".cond.bias", {2 * W}});
  v.push_back({"y_cat_emb.weight", {c.n_types, E}});
  lin("y_cont_mlp.0", c.y_cont_dim, E); lin("y_cont_mlp.2", E, E);
  lin("y_fuse.0", 2 * E, W); lin("y_fuse.2", W, W);
  lin("t_mlp.0", c.t_emb_dim, W); lin("t_mlp.2", W, W);
  lin("in_proj", c.z_dim, W);
  for (int i = 0; i < c.n_blocks; ++i) {
    const std::string p = "blocks." + std::to_string(i) + ".";
    ln(p + "norm"); lin(p + "fc1", W, 4 * W); lin(p + "fc2", 4 * W, W);
  }
  ln("out_norm"); lin("out_proj", W, c.z_dim);
  return v;
}

static int prior_pinned(tcs_prior* h, size_t bytes) {
  TCS_CUDA(cudaEventSynchronize(h->ev_pinned));
  if (h->pinned_bytes >= bytes) return TCS_OK;
  if (h->pinned) cudaFreeHost(h->pinned);
  h->pinned = nullptr; h->pinned_bytes = 0;
  TCS_CUDA(cudaMallocHost(&h->pinned, bytes));
  h->pinned_bytes = bytes;
  return TCS_OK;
}

// workspace for `n` rows and `s` step rows
static int prior_workspace(tcs_prior* h, int n, int s) {
  if (n <= h->cap_n && s <= h->cap_s) return TCS_OK;
  const size_t N = n > h->cap_n ? n : h->cap_n, S = s > h->cap_s ? s : h->cap_s;
  const size_t W = h->cfg.width, B = h->cfg.n_blocks, E = h->cfg.y_cat_emb_dim, esz = h->bf16 ? 2 : 4;
  TCS_CHECK(h->h.ensure(N * W * 4));
  TCS_CHECK(h->u.ensure(N * W * esz));
  TCS_CHECK(h->a.ensure(N * 4 * W * esz));
  TCS_CHECK(h->z.ensure(N * h->cfg.z_dim * 4));
  TCS_CHECK(h->film.ensure(N * B * 2 * W * 4));
  TCS_CHECK(h->ycat.ensure(N * 2 * E * 4));
  TCS_CHECK(h->y1.ensure(N * W * esz));
  TCS_CHECK(h->yfeat.ensure(N * W * esz));
  TCS_CHECK(h->condin.ensure(N * 2 * W * esz));
  const size_t R = N > S ? N : S;   // the t-chain runs over S rows (sampling) or n rows (tcs_prior_eps)
  TCS_CHECK(h->te.ensure(R * h->cfg.t_emb_dim * 4));
  TCS_CHECK(h->t1.ensure(R * W * 4));
  TCS_CHECK(h->tfeat.ensure(S * W * 4));
  TCS_CHECK(h->tcond.ensure(S * B * 2 * W * 4));
  TCS_CHECK(h->coef.ensure(S * sizeof(DdimCoef)));
  TCS_CHECK(h->ts32.ensure(S * 4));
  h->cap_n = static_cast<int>(N); h->cap_s = static_cast<int>(S);
  ++h->ws_gen;
  return TCS_OK;
}

// out = act(A W^T + bias): tcgen05 when `tc` (bf16 A and W), FFMA otherwise (fp32 A and W)
static int dense(tcs_prior* h, bool tc, const void* A, int lda, int ldw, const float* w32, const __nv_bfloat16* w16, int M,
                 int N, int K, const float* bias, void* out, int ldo, int flags, cudaStream_t st) {
  ++h->launches;
  if (tc) {
    LinearTcPlan pl;
    TCS_CHECK(linear_tc_make_plan(&pl, static_cast<const __nv_bfloat16*>(A), lda, w16, ldw, M, N, K, bias, out, ldo,
                                  flags, h->sm_count));
    return linear_tc_launch(pl, st);
  }
  return launch_linear_simt(static_cast<const float*>(A), lda, w32, ldw, M, N, K, bias, out, ldo, flags, st);
}

// y-chain: ycat -> y1 -> y_feat, written to `yfeat_out` (row pitch ld; bf16 in bf16 mode, fp32 otherwise)
static int y_chain(tcs_prior* h, const int64_t* y_cat, const float* y_cont, int n, void* yfeat_out, int ld, cudaStream_t st) {
  const int W = h->cfg.width, E = h->cfg.y_cat_emb_dim;
  ++h->launches;
  TCS_CHECK(launch_prior_ycat(h->ew, y_cat, y_cont, n, h->ycat.as<float>(), st));
  const int of = h->bf16 ? 0 : LIN_OUT_F32;
  TCS_CHECK(dense(h, false, h->ycat.p, 2 * E, 2 * E, h->dw.at("y_fuse.0.weight"), nullptr, n, W, 2 * E,
                  h->dw.at("y_fuse.0.bias"), h->y1.p, W, LIN_SILU | of, st));
  TCS_CHECK(dense(h, h->bf16, h->y1.p, W, W, h->dw.at("y_fuse.2.weight"), h->bf16 ? h->dw16.at("y_fuse.2.weight") : nullptr,
                  n, W, W, h->dw.at("y_fuse.2.bias"), yfeat_out, ld, of, st));
  return TCS_OK;
}

// the FiLM blocks over h (fp32 [n, W]); film_row [n, B*2W] and/or film_step [S, B*2W]
static int run_blocks(tcs_prior* h, int n, const void* film_row, const float* film_step, const int* step_ptr, cudaStream_t st) {
  const int W = h->cfg.width, B = h->cfg.n_blocks, fld = B * 2 * W;
  for (int b = 0; b < B; ++b) {
    const std::string p = "blocks." + std::to_string(b) + ".";
    ++h->launches;
    if (h->bf16)
      TCS_CHECK(launch_ln_film<__nv_bfloat16>(h->h.as<float>(), n, W, h->dw.at(p + "norm.weight"), h->dw.at(p + "norm.bias"),
                                              static_cast<const __nv_bfloat16*>(film_row), fld, film_step, fld, step_ptr, b * 2 * W,
                                              h->u.as<__nv_bfloat16>(), st));
    else
      TCS_CHECK(launch_ln_film<float>(h->h.as<float>(), n, W, h->dw.at(p + "norm.weight"), h->dw.at(p + "norm.bias"),
                                      static_cast<const float*>(film_row), fld, film_step, fld, step_ptr, b * 2 * W, h->u.as<float>(), st));
    const int of = h->bf16 ? 0 : LIN_OUT_F32;
    TCS_CHECK(dense(h, h->bf16, h->u.p, W, W, h->dw.at(p + "fc1.weight"), h->bf16 ? h->dw16.at(p + "fc1.weight") : nullptr,
                    n, 4 * W, W, h->dw.at(p + "fc1.bias"), h->a.p, 4 * W, LIN_SILU | of, st));
    TCS_CHECK(dense(h, h->bf16, h->a.p, 4 * W, 4 * W, h->dw.at(p + "fc2.weight"),
                    h->bf16 ? h->dw16.at(p + "fc2.weight") : nullptr, n, W, 4 * W, h->dw.at(p + "fc2.bias"), h->h.p, W,
                    LIN_OUT_F32 | LIN_ACCUM, st));
  }
  return TCS_OK;
}

static PriorTailArgs tail_args(tcs_prior* h, int mode, int n) {
  PriorTailArgs a{};
  a.mode = mode; a.n = n; a.W = h->cfg.width; a.zd = h->cfg.z_dim;
  a.h = h->h.as<float>(); a.z = h->z.as<float>();
  a.on_w = h->dw.at("out_norm.weight"); a.on_b = h->dw.at("out_norm.bias");
  a.op_w = h->dw.at("out_proj.weight"); a.op_b = h->dw.at("out_proj.bias");
  a.ip_w = h->dw.at("in_proj.weight"); a.ip_b = h->dw.at("in_proj.bias");
  a.coef = h->coef.as<DdimCoef>(); a.step_ptr = h->step_ctr.as<int>();
  return a;
}

static int check_prior(tcs_prior* h, const char* who) {
  if (!h) return fail(TCS_ERR_BAD_ARGUMENT, std::string(who) + ": null handle");
  if (!h->finalized) return fail(TCS_ERR_STATE, std::string(who) + ": call tcs_prior_finalize_weights first");
  return TCS_OK;
}

// one DDIM run over a chunk of samples
static int ddim_chunk(tcs_prior* h, const tcs_ddim_args& A, int row0, int n, const std::vector<DdimCoef>& coef,
                      const std::vector<int64_t>& ts, cudaStream_t st) {
  const int W = h->cfg.width, B = h->cfg.n_blocks, zd = h->cfg.z_dim, S = static_cast<int>(ts.size());
  const int fld = B * 2 * W;
  TCS_CHECK(prior_workspace(h, n, S));
  // ---- tables -----------------------------------------------------------------------------------------
  const size_t pin = S * sizeof(DdimCoef) + S * 4 + 64;
  TCS_CHECK(prior_pinned(h, pin));
  DdimCoef* hc = static_cast<DdimCoef*>(h->pinned);
  int* hts = reinterpret_cast<int*>(hc + S);
  for (int i = 0; i < S; ++i) { hc[i] = coef[i]; hts[i] = static_cast<int>(ts[i]); }
  TCS_CUDA(cudaMemcpyAsync(h->coef.p, hc, S * sizeof(DdimCoef), cudaMemcpyHostToDevice, st));
  TCS_CUDA(cudaMemcpyAsync(h->ts32.p, hts, S * 4, cudaMemcpyHostToDevice, st));
  TCS_CUDA(cudaEventRecord(h->ev_pinned, st));
  TCS_CUDA(cudaMemsetAsync(h->step_ctr.p, 0, 4, st));
  // ---- t-chain (S rows, fp32) ----------------------------------------------------------------------------
  ++h->launches;
  TCS_CHECK(launch_prior_time_features(h->ts32.p, 1, h->freqs, S, h->cfg.t_emb_dim, h->te.as<float>(), st));
  TCS_CHECK(dense(h, false, h->te.p, h->cfg.t_emb_dim, h->cfg.t_emb_dim, h->dw.at("t_mlp.0.weight"), nullptr, S, W,
                  h->cfg.t_emb_dim, h->dw.at("t_mlp.0.bias"), h->t1.p, W, LIN_SILU | LIN_OUT_F32, st));
  TCS_CHECK(dense(h, false, h->t1.p, W, W, h->dw.at("t_mlp.2.weight"), nullptr, S, W, W, h->dw.at("t_mlp.2.bias"),
                  h->tfeat.p, W, LIN_OUT_F32, st));
  TCS_CHECK(dense(h, false, h->tfeat.p, W, 2 * W, h->cond_w, nullptr, S, fld, W, h->cond_b, h->tcond.p, fld, LIN_OUT_F32, st));
  // ---- y-chain and the step-invariant half of every FiLM projection ------------------------------------------
  TCS_CHECK(y_chain(h, A.y_cat + row0, A.y_cont + static_cast<size_t>(row0) * h->cfg.y_cont_dim, n, h->yfeat.p, W, st));
  TCS_CHECK(dense(h, h->bf16, h->yfeat.p, W, 2 * W, h->cond_w + W, h->bf16 ? h->cond_w16 + W : nullptr, n, fld, W, nullptr,
                  h->film.p, fld, h->bf16 ? 0 : LIN_OUT_F32, st));
  // ---- initial state ------------------------------------------------------------------------------------------
  ++h->launches;
  if (A.z_init)
    TCS_CHECK(launch_copy_f32(h->z.as<float>(), A.z_init + static_cast<size_t>(row0) * zd, static_cast<size_t>(n) * zd, st));
  else
    TCS_CHECK(launch_prior_philox(h->z.as<float>(), n, zd, A.seed, A.global_index_offset + row0, st));
  ++h->launches;
  TCS_CHECK(launch_prior_tail(tail_args(h, TAIL_INIT, n), st));

  auto enqueue_step = [&]() -> int {
    TCS_CHECK(run_blocks(h, n, h->film.p, h->tcond.as<float>(), h->step_ctr.as<int>(), st));
    PriorTailArgs ta = tail_args(h, TAIL_DDIM, n);
    ta.z_out = A.z0_out + static_cast<size_t>(row0) * zd;
    // traces are [S, n_total, zd]: this chunk's rows start at row0
    ta.trace_eps = A.trace_eps ? A.trace_eps + static_cast<size_t>(row0) * zd : nullptr;
    ta.trace_z = A.trace_z ? A.trace_z + static_cast<size_t>(row0) * zd : nullptr;
    ta.trace_n = A.n;
    ++h->launches;
    TCS_CHECK(launch_prior_tail(ta, st));
    ++h->launches;
    TCS_CHECK(launch_advance(h->step_ctr.as<int>(), 1, st));
    return TCS_OK;
  };

  if (h->cfg.use_graph) {
    tcs_prior::GKey k;
    k.n = n; k.gen = h->ws_gen; k.z0 = A.z0_out + static_cast<size_t>(row0) * zd; k.teps = A.trace_eps; k.tz = A.trace_z;
    k.tn = A.n; k.trow0 = row0;
    const bool same = h->gexec && h->gkey.n == k.n && h->gkey.gen == k.gen && h->gkey.z0 == k.z0 && h->gkey.teps == k.teps &&
                      h->gkey.tz == k.tz && h->gkey.tn == k.tn && h->gkey.trow0 == k.trow0;
    if (!same) {
      if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
      const int64_t before = h->launches;
      TCS_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      const int rc = enqueue_step();
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rc != TCS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess) return fail(TCS_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
      h->graph_kernels = h->launches - before;
      h->launches = before;
      const cudaError_t ie = cudaGraphInstantiate(&h->gexec, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) { h->gexec = nullptr; return fail(TCS_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
      h->gkey = k;
    }
    for (int i = 0; i < S; ++i) TCS_CUDA(cudaGraphLaunch(h->gexec, st));
    h->launches += h->graph_kernels * S;
  } else {
    for (int i = 0; i < S; ++i) TCS_CHECK(enqueue_step());
  }
  return TCS_OK;
}

}  // namespace tcs

extern "C" {

void tcs_prior_default_config(tcs_prior_config* c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->z_dim = 32; c->n_types = 4; c->y_cont_dim = 4; c->t_emb_dim = 64; c->width = 1024; c->n_blocks = 8; c->y_cat_emb_dim = 64;
  c->T = 1000; c->beta_start = 1e-4; c->beta_end = 0.05;
  c->precision = TCS_BF16; c->device = 0; c->use_graph = 1;
}

int tcs_prior_schedule_host(int32_t T, double beta_start, double beta_end, float* out) {
  if (T < 1 || !out) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_schedule_host: T >= 1 and an output buffer are required");
  std::vector<float> ab;
  alpha_bars_host(T, beta_start, beta_end, &ab);
  memcpy(out, ab.data(), sizeof(float) * T);
  return TCS_OK;
}
int tcs_prior_timesteps_host(int32_t T, int32_t n_steps, int64_t* ts_out, int32_t* count) {
  if (T < 1 || n_steps < 1 || !ts_out || !count) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_timesteps_host: bad argument");
  std::vector<int64_t> ts;
  ddim_timesteps_host(T, n_steps, &ts);
  memcpy(ts_out, ts.data(), sizeof(int64_t) * ts.size());
  *count = static_cast<int32_t>(ts.size());
  return TCS_OK;
}
int64_t tcs_prior_launch_count(const tcs_prior* h) { return h ? h->launches : 0; }
int64_t tcs_vae_launch_count(const tcs_vae* h) { return h ? h->launches : 0; }

int tcs_prior_create(tcs_prior** out, const tcs_prior_config* cfg) {
  if (!out || !cfg) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_create: null argument");
  *out = nullptr;
  const tcs_prior_config& c = *cfg;
  if (c.width != 256 && c.width != 512 && c.width != 1024 && c.width != 2048)
    return fail(TCS_ERR_UNSUPPORTED, "this build supports width 256, 512, 1024 or 2048 (got " + std::to_string(c.width) + ")");
  if (c.z_dim < 4 || c.z_dim > PRIOR_MAX_Z || c.z_dim % 4)
    return fail(TCS_ERR_UNSUPPORTED, "this build supports z_dim in {4, 8, ..., 32} (got " + std::to_string(c.z_dim) + ")");
  if (c.t_emb_dim < 16 || c.t_emb_dim % 16) return fail(TCS_ERR_UNSUPPORTED, "t_emb_dim must be a positive multiple of 16");
  if (c.y_cat_emb_dim != 32 && c.y_cat_emb_dim != 64 && c.y_cat_emb_dim != 128)
    return fail(TCS_ERR_UNSUPPORTED, "y_cat_emb_dim must be 32, 64 or 128");
  if (c.n_types < 1 || c.y_cont_dim < 1 || c.y_cont_dim > 16 || c.n_blocks < 1 || c.n_blocks > 64)
    return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_create: need n_types >= 1, 1 <= y_cont_dim <= 16, 1 <= n_blocks <= 64");
  if (c.T < 1 || !(c.beta_start > 0.0) || !(c.beta_end < 1.0 + 1e-12) || c.beta_end < c.beta_start)
    return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_create: need T >= 1 and 0 < beta_start <= beta_end <= 1");
  if (c.precision != TCS_FP32 && c.precision != TCS_BF16) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_create: bad precision");
  TCS_CHECK(check_device(c.device));
  std::unique_ptr<tcs_prior> h(new tcs_prior());
  h->cfg = c;
  h->bf16 = c.precision == TCS_BF16;
  cudaDeviceProp prop;
  TCS_CUDA(cudaGetDeviceProperties(&prop, c.device));
  h->sm_count = prop.multiProcessorCount;
  TCS_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  TCS_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
  TCS_CUDA(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
  TCS_CUDA(cudaEventCreateWithFlags(&h->ev_pinned, cudaEventDisableTiming));
  TCS_CHECK(h->step_ctr.ensure(16));
  *out = h.release();
  return TCS_OK;
}

void tcs_prior_destroy(tcs_prior* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  delete h;
}

int tcs_prior_set_weight(tcs_prior* h, const char* key, const float* data, const int64_t* shape, int32_t ndim) {
  if (!h || !key || !data || !shape || ndim < 1 || ndim > 4) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_set_weight: bad argument");
  TCS_CUDA(cudaSetDevice(h->cfg.device));
  h->finalized = false;
  return store_weight(&h->host_w, key, data, shape, ndim);
}

int tcs_prior_finalize_weights(tcs_prior* h) {
  if (!h) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_finalize_weights: null handle");
  TCS_CUDA(cudaSetDevice(h->cfg.device));
  const tcs_prior_config& c = h->cfg;
  const Expected exp = prior_expected(c);
  size_t total = 0;
  for (const auto& kv : exp) {
    auto it = h->host_w.find(kv.first);
    if (it == h->host_w.end()) return fail(TCS_ERR_STATE, "missing state-dict tensor: " + kv.first);
    if (it->second.shape != kv.second) return fail(TCS_ERR_BAD_ARGUMENT, "state-dict tensor has the wrong shape: " + kv.first);
    total += (it->second.v.size() + 63) / 64 * 64;
  }
  if (h->host_w.size() != exp.size()) return fail(TCS_ERR_BAD_ARGUMENT, "unexpected extra state-dict tensors");
  const int half = c.t_emb_dim / 2;
  TCS_CHECK(h->arena.ensure((total + 64 + half) * 4));
  float* base = h->arena.as<float>();
  size_t off = 0;
  h->dw.clear();
  std::map<std::string, size_t> offs;
  for (const auto& kv : exp) {
    const HostTensor& t = h->host_w.at(kv.first);
    TCS_CUDA(cudaMemcpy(base + off, t.v.data(), t.v.size() * 4, cudaMemcpyHostToDevice));
    h->dw[kv.first] = base + off;
    offs[kv.first] = off;
    off += (t.v.size() + 63) / 64 * 64;   // (2W)^2 and 2W are multiples of 64: the cond tensors stay contiguous
  }
  h->cond_w = h->dw.at("blocks.0.cond.weight");
  h->cond_b = h->dw.at("blocks.0.cond.bias");
  {  // timestep_embedding frequencies: exp(-linspace(0, ln 1e4, half)) in fp32 (:18-21)
    std::vector<float> lin, fr(half);
    linspace_f32(0.0f, static_cast<float>(std::log(10000.0)), half, &lin);
    for (int i = 0; i < half; ++i) fr[i] = expf(lin[i] * -1.0f);
    TCS_CUDA(cudaMemcpy(base + off, fr.data(), half * 4, cudaMemcpyHostToDevice));
    h->freqs = base + off;
  }
  PriorEmbedWeights& ew = h->ew;
  ew.cat_emb = h->dw.at("y_cat_emb.weight");
  ew.cm0_w = h->dw.at("y_cont_mlp.0.weight"); ew.cm0_b = h->dw.at("y_cont_mlp.0.bias");
  ew.cm2_w = h->dw.at("y_cont_mlp.2.weight"); ew.cm2_b = h->dw.at("y_cont_mlp.2.bias");
  ew.n_types = c.n_types; ew.y_cont_dim = c.y_cont_dim; ew.E = c.y_cat_emb_dim;
  h->dw16.clear();
  if (h->bf16) {   // bf16 copies of the operands of the tensor-core GEMMs (same offsets as the fp32 arena)
    TCS_CHECK(h->arena16.ensure(total * 2));
    TCS_CHECK(launch_f32_to_bf16(base, h->arena16.as<__nv_bfloat16>(), total, h->stream));
    TCS_CUDA(cudaStreamSynchronize(h->stream));
    for (const auto& kv : offs) h->dw16[kv.first] = h->arena16.as<__nv_bfloat16>() + kv.second;
    h->cond_w16 = h->dw16.at("blocks.0.cond.weight");
  }
  if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; h->gkey.n = -1; }
  h->finalized = true;
  return TCS_OK;
}

int tcs_prior_eps(tcs_prior* h, const float* z_t, const int64_t* t, const int64_t* y_cat, const float* y_cont, int32_t n,
                  float* eps_out, void* stream) {
  TCS_CHECK(check_prior(h, "tcs_prior_eps"));
  if (n < 0 || (n > 0 && (!z_t || !t || !y_cat || !y_cont || !eps_out))) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_eps: null tensor");
  if (n == 0) return TCS_OK;
  cudaStream_t user = static_cast<cudaStream_t>(stream);
  TCS_CHECK(enter_h(h, user));
  cudaStream_t st = h->stream;
  const int W = h->cfg.width, B = h->cfg.n_blocks, zd = h->cfg.z_dim, fld = B * 2 * W, temb = h->cfg.t_emb_dim;
  const size_t esz = h->bf16 ? 2 : 4;
  for (int row0 = 0; row0 < n; row0 += PRIOR_CHUNK) {
    const int m = n - row0 < PRIOR_CHUNK ? n - row0 : PRIOR_CHUNK;
    TCS_CHECK(prior_workspace(h, m, 1));
    // cond = [t_feat, y_feat] (:116), built in place: columns [0, W) and [W, 2W) of condin
    ++h->launches;
    TCS_CHECK(launch_prior_time_features(t + row0, 0, h->freqs, m, temb, h->te.as<float>(), st));
    const int of = h->bf16 ? 0 : LIN_OUT_F32;
    // t1 lives in `a` here (free until the blocks run)
    TCS_CHECK(dense(h, false, h->te.p, temb, temb, h->dw.at("t_mlp.0.weight"), nullptr, m, W, temb, h->dw.at("t_mlp.0.bias"),
                    h->a.p, W, LIN_SILU | of, st));
    TCS_CHECK(dense(h, h->bf16, h->a.p, W, W, h->dw.at("t_mlp.2.weight"), h->bf16 ? h->dw16.at("t_mlp.2.weight") : nullptr, m,
                    W, W, h->dw.at("t_mlp.2.bias"), h->condin.p, 2 * W, of, st));
    TCS_CHECK(y_chain(h, y_cat + row0, y_cont + static_cast<size_t>(row0) * h->cfg.y_cont_dim, m,
                      static_cast<uint8_t*>(h->condin.p) + static_cast<size_t>(W) * esz, 2 * W, st));
    TCS_CHECK(dense(h, h->bf16, h->condin.p, 2 * W, 2 * W, h->cond_w, h->cond_w16, m, fld, 2 * W, h->cond_b, h->film.p, fld,
                    h->bf16 ? 0 : LIN_OUT_F32, st));
    PriorTailArgs ia = tail_args(h, TAIL_INIT, m);
    ia.z = const_cast<float*>(z_t) + static_cast<size_t>(row0) * zd;
    ++h->launches;
    TCS_CHECK(launch_prior_tail(ia, st));
    TCS_CHECK(run_blocks(h, m, h->film.p, nullptr, nullptr, st));
    PriorTailArgs ea = tail_args(h, TAIL_EPS, m);
    ea.eps_out = eps_out + static_cast<size_t>(row0) * zd;
    ++h->launches;
    TCS_CHECK(launch_prior_tail(ea, st));
  }
  return leave_h(h, user);
}

int tcs_prior_ddim_sample(tcs_prior* h, const tcs_ddim_args* args, void* stream) {
  TCS_CHECK(check_prior(h, "tcs_prior_ddim_sample"));
  if (!args) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_ddim_sample: null args");
  const tcs_ddim_args& A = *args;
  if (A.n < 0 || A.n_steps < 1) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_ddim_sample: need n >= 0 and n_steps >= 1");
  if (A.n == 0) return TCS_OK;
  if (!A.y_cat || !A.y_cont || !A.z0_out) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_ddim_sample: y_cat, y_cont and z0_out are required");
  std::vector<float> abar;
  alpha_bars_host(h->cfg.T, h->cfg.beta_start, h->cfg.beta_end, &abar);
  std::vector<int64_t> ts;
  ddim_timesteps_host(h->cfg.T, A.n_steps, &ts);
  const int S = static_cast<int>(ts.size());
  std::vector<DdimCoef> coef(S);
  for (int i = 0; i < S; ++i) {
    DdimCoef c{};
    const float ab = abar[ts[i]];
    c.s1m_t = sqrtf(1.0f - ab);
    c.sa_t_eps = sqrtf(ab) + 1e-8f;
    c.last = i == S - 1;
    c.t = static_cast<int>(ts[i]);
    if (!c.last) {
      const float abp = abar[ts[i + 1]];
      c.sa_prev = sqrtf(abp);
      c.s1m_prev = sqrtf(1.0f - abp);
    }
    coef[i] = c;
  }
  cudaStream_t user = static_cast<cudaStream_t>(stream);
  TCS_CHECK(enter_h(h, user));
  for (int row0 = 0; row0 < A.n; row0 += PRIOR_CHUNK) {
    const int m = A.n - row0 < PRIOR_CHUNK ? A.n - row0 : PRIOR_CHUNK;
    TCS_CHECK(ddim_chunk(h, A, row0, m, coef, ts, h->stream));
  }
  return leave_h(h, user);
}

int tcs_prior_profile(tcs_prior* h, int32_t n, int32_t reps, float* ms) {
  TCS_CHECK(check_prior(h, "tcs_prior_profile"));
  if (n < 1 || reps < 1 || !ms) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_prior_profile: bad argument");
  if (n > h->cap_n || h->cap_s < 1) return fail(TCS_ERR_STATE, "tcs_prior_profile: run tcs_prior_ddim_sample on >= n samples first");
  TCS_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t st = h->stream;
  const int W = h->cfg.width, fld = h->cfg.n_blocks * 2 * W, of = h->bf16 ? 0 : LIN_OUT_F32;
  cudaEvent_t ev[5];
  for (auto& e : ev) TCS_CUDA(cudaEventCreate(&e));
  TCS_CUDA(cudaMemsetAsync(h->step_ctr.p, 0, 4, st));
  const std::string p = "blocks.0.";
  auto fc1 = [&]() {
    return dense(h, h->bf16, h->u.p, W, W, h->dw.at(p + "fc1.weight"), h->bf16 ? h->dw16.at(p + "fc1.weight") : nullptr, n, 4 * W, W,
                 h->dw.at(p + "fc1.bias"), h->a.p, 4 * W, LIN_SILU | of, st);
  };
  auto fc2 = [&]() {
    return dense(h, h->bf16, h->a.p, 4 * W, 4 * W, h->dw.at(p + "fc2.weight"), h->bf16 ? h->dw16.at(p + "fc2.weight") : nullptr, n, W,
                 4 * W, h->dw.at(p + "fc2.bias"), h->h.p, W, LIN_OUT_F32 | LIN_ACCUM, st);
  };
  auto lnf = [&]() {
    if (h->bf16)
      return launch_ln_film<__nv_bfloat16>(h->h.as<float>(), n, W, h->dw.at(p + "norm.weight"), h->dw.at(p + "norm.bias"),
                                           h->film.as<__nv_bfloat16>(), fld, h->tcond.as<float>(), fld, h->step_ctr.as<int>(), 0,
                                           h->u.as<__nv_bfloat16>(), st);
    return launch_ln_film<float>(h->h.as<float>(), n, W, h->dw.at(p + "norm.weight"), h->dw.at(p + "norm.bias"),
                                 h->film.as<float>(), fld, h->tcond.as<float>(), fld, h->step_ctr.as<int>(), 0, h->u.as<float>(), st);
  };
  auto tail = [&]() {
    PriorTailArgs ta = tail_args(h, TAIL_DDIM, n);
    ta.z_out = h->z.as<float>();
    ta.trace_n = n;
    return launch_prior_tail(ta, st);
  };
  TCS_CHECK(launch_prior_tail(tail_args(h, TAIL_INIT, n), st));   // a sane residual stream from the current latents
  TCS_CHECK(lnf()); TCS_CHECK(fc1()); TCS_CHECK(fc2()); TCS_CHECK(tail());   // warm-up
  TCS_CUDA(cudaEventRecord(ev[0], st));
  for (int r = 0; r < reps; ++r) TCS_CHECK(fc1());
  TCS_CUDA(cudaEventRecord(ev[1], st));
  for (int r = 0; r < reps; ++r) TCS_CHECK(fc2());
  TCS_CUDA(cudaEventRecord(ev[2], st));
  for (int r = 0; r < reps; ++r) TCS_CHECK(lnf());
  TCS_CUDA(cudaEventRecord(ev[3], st));
  for (int r = 0; r < reps; ++r) TCS_CHECK(tail());
  TCS_CUDA(cudaEventRecord(ev[4], st));
  TCS_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 4; ++k) {
    TCS_CUDA(cudaEventElapsedTime(ms + k, ev[k], ev[k + 1]));
    ms[k] /= static_cast<float>(reps);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return TCS_OK;
}

// ---- CondVAE decoder --------------------------------------------------------------------------------------
int tcs_vae_create(tcs_vae** out, const tcs_vae_config* cfg) {
  if (!out || !cfg) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_vae_create: null argument");
  *out = nullptr;
  if (cfg->z_dim < 1 || cfg->z_dim > 32 || cfg->n_types < 1 || cfg->y_cont_dim < 0 || cfg->y_cont_dim > 16)
    return fail(TCS_ERR_UNSUPPORTED, "tcs_vae_create: need 1 <= z_dim <= 32, n_types >= 1, 0 <= y_cont_dim <= 16");
  TCS_CHECK(check_device(cfg->device));
  if (cfg->precision != TCS_FP32 && cfg->precision != TCS_BF16) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_vae_create: bad precision");
  std::unique_ptr<tcs_vae> h(new tcs_vae());
  h->cfg = *cfg;
  h->bf16 = cfg->precision == TCS_BF16;
  {
    cudaDeviceProp prop;
    TCS_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;
  }
  TCS_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  TCS_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
  TCS_CUDA(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
  *out = h.release();
  return TCS_OK;
}
void tcs_vae_destroy(tcs_vae* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  delete h;
}
int tcs_vae_set_weight(tcs_vae* h, const char* key, const float* data, const int64_t* shape, int32_t ndim) {
  if (!h || !key || !data || !shape || ndim < 1 || ndim > 4) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_vae_set_weight: bad argument");
  TCS_CUDA(cudaSetDevice(h->cfg.device));
  h->finalized = false;
  const std::string k(key);
  if (k.rfind("enc.", 0) == 0 || k.rfind("enc_fc.", 0) == 0 || k.rfind("mu.", 0) == 0 || k.rfind("logvar.", 0) == 0)
    return TCS_OK;   // encoder: not on the sampling path
  return store_weight(&h->host_w, key, data, shape, ndim);
}
int tcs_vae_finalize_weights(tcs_vae* h) {
  if (!h) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_vae_finalize_weights: null handle");
  TCS_CUDA(cudaSetDevice(h->cfg.device));
  const int64_t in_dim = h->cfg.z_dim + h->cfg.n_types + h->cfg.y_cont_dim;
  static const int dch[5] = {256, 128, 64, 32, 1};
  Expected exp = {{"dec_fc.weight", {4096, in_dim}}, {"dec_fc.bias", {4096}}};
  for (int i = 0; i < 4; ++i) {
    exp.push_back({"dec." + std::to_string(2 * i) + ".weight", {dch[i], dch[i + 1], 4, 4}});
    exp.push_back({"dec." + std::to_string(2 * i) + ".bias", {dch[i + 1]}});
  }
  for (const auto& kv : exp) {
    auto it = h->host_w.find(kv.first);
    if (it == h->host_w.end()) return fail(TCS_ERR_STATE, "missing state-dict tensor: " + kv.first);
    if (it->second.shape != kv.second) return fail(TCS_ERR_BAD_ARGUMENT, "state-dict tensor has the wrong shape: " + kv.first);
  }
  if (h->host_w.size() != exp.size()) return fail(TCS_ERR_BAD_ARGUMENT, "unexpected extra state-dict tensors");
  std::vector<float> blob;
  auto put = [&](const std::vector<float>& v) { const size_t o = blob.size(); blob.insert(blob.end(), v.begin(), v.end()); blob.resize((blob.size() + 63) / 64 * 64); return o; };
  const size_t o_fcw = put(h->host_w.at("dec_fc.weight").v), o_fcb = put(h->host_w.at("dec_fc.bias").v);
  size_t o_w[4], o_b[3];
  for (int i = 0; i < 4; ++i) {
    const HostTensor& w = h->host_w.at("dec." + std::to_string(2 * i) + ".weight");
    std::vector<float> pk(w.v.size());
    vae_convt_pack_weights(w.v.data(), dch[i], dch[i + 1], pk.data());
    o_w[i] = put(pk);
    if (i < 3) o_b[i] = put(h->host_w.at("dec." + std::to_string(2 * i) + ".bias").v);
  }
  h->out_bias = h->host_w.at("dec.6.bias").v[0];
  TCS_CHECK(h->arena.ensure(blob.size() * 4));
  TCS_CUDA(cudaMemcpy(h->arena.p, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice));
  const float* base = h->arena.as<float>();
  h->fc_w = base + o_fcw; h->fc_b = base + o_fcb;
  for (int i = 0; i < 4; ++i) h->ct_w[i] = base + o_w[i];
  for (int i = 0; i < 3; ++i) h->ct_b[i] = base + o_b[i];
  if (h->bf16) {   // bf16 copies of the packed weights of the three tensor-core stages
    TCS_CHECK(h->arena16.ensure(blob.size() * 2));
    TCS_CHECK(launch_f32_to_bf16(base, h->arena16.as<__nv_bfloat16>(), blob.size(), h->stream));
    TCS_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 3; ++i) h->ct_w16[i] = h->arena16.as<__nv_bfloat16>() + o_w[i];
  }
  h->finalized = true;
  return TCS_OK;
}

int tcs_vae_decode(tcs_vae* h, const float* z, const int64_t* y_cat, const float* y_cont, int32_t n, const float* z_mean,
                   const float* z_std, float* x_out, void* stream) {
  if (!h) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_vae_decode: null handle");
  if (!h->finalized) return fail(TCS_ERR_STATE, "tcs_vae_decode: call tcs_vae_finalize_weights first");
  if (n < 0 || (n > 0 && (!z || !y_cat || (!y_cont && h->cfg.y_cont_dim > 0) || !x_out)))
    return fail(TCS_ERR_BAD_ARGUMENT, "tcs_vae_decode: null tensor");
  if ((z_mean == nullptr) != (z_std == nullptr)) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_vae_decode: give both z_mean and z_std or neither");
  if (n == 0) return TCS_OK;
  cudaStream_t user = static_cast<cudaStream_t>(stream);
  TCS_CHECK(enter_h(h, user));
  cudaStream_t st = h->stream;
  const int zd = h->cfg.z_dim, ycd = h->cfg.y_cont_dim;
  const size_t cap = n < VAE_CHUNK ? n : VAE_CHUNK;
  const size_t esz = h->bf16 ? 2 : 4;
  TCS_CHECK(h->a0.ensure(cap * 16 * 256 * esz));
  TCS_CHECK(h->a1.ensure(cap * 64 * 128 * esz));
  TCS_CHECK(h->a2.ensure(cap * 256 * 64 * esz));
  TCS_CHECK(h->a3.ensure(cap * 1024 * 32 * esz));
  static const int kHi[3] = {4, 8, 16}, kCi[3] = {256, 128, 64}, kCo[3] = {128, 64, 32};
  DevBuf* act[4] = {&h->a0, &h->a1, &h->a2, &h->a3};
  for (int i0 = 0; i0 < n; i0 += VAE_CHUNK) {
    const int m = n - i0 < VAE_CHUNK ? n - i0 : VAE_CHUNK;
    h->launches += 5;
    TCS_CHECK(launch_vae_dec_fc(z + static_cast<size_t>(i0) * zd, y_cat + i0, y_cont + static_cast<size_t>(i0) * ycd, z_mean, z_std,
                                h->fc_w, h->fc_b, m, zd, h->cfg.n_types, ycd, h->a0.p, h->bf16 ? 1 : 0, st));
    for (int l = 0; l < 3; ++l) {
      if (h->bf16) {   // tcgen05: the four parity classes of the layer as GEMMs with K = 4 taps x C_in
        LinearTcPlan pl;
        TCS_CHECK(linear_tc_make_convt_plan(&pl, act[l]->as<__nv_bfloat16>(), h->ct_w16[l], h->ct_b[l], m, kHi[l], kCi[l], kCo[l],
                                            act[l + 1]->as<__nv_bfloat16>(), h->sm_count));
        TCS_CHECK(linear_tc_launch(pl, st));
      } else {
        TCS_CHECK(launch_vae_convt(act[l]->as<float>(), h->ct_w[l], h->ct_b[l], m, kHi[l], kCi[l], kCo[l], act[l + 1]->as<float>(), st));
      }
    }
    TCS_CHECK(launch_vae_convt_out(h->a3.p, h->bf16 ? 1 : 0, h->ct_w[3], h->out_bias, m, x_out + static_cast<size_t>(i0) * 4096, st));
  }
  return leave_h(h, user);
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, size_t count) {
  const size_t i = blockIdx.x * 256ULL + threadIdx.x;
  if (i < count) dst[i] = __bfloat162float(src[i]);
}

int tcs_debug_linear(int32_t engine, int32_t M, int32_t N, int32_t K, const float* A, const float* W, const float* bias,
                     float* out, int32_t silu, int32_t accumulate, int32_t bf16_out, void* stream) {
  if (!A || !W || !out || M < 1 || N < 1 || K < 1) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_debug_linear: bad argument");
  if (accumulate && bf16_out) return fail(TCS_ERR_BAD_ARGUMENT, "tcs_debug_linear: accumulate needs fp32 output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int flags = (silu ? LIN_SILU : 0) | (bf16_out ? 0 : LIN_OUT_F32) | (accumulate ? LIN_ACCUM : 0);
  DevBuf a16, w16, o16;
  const size_t mn = static_cast<size_t>(M) * N;
  void* dst = out;
  if (bf16_out) { TCS_CHECK(o16.ensure(mn * 2)); dst = o16.p; }
  if (engine == TCS_ENGINE_TCGEN05) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    TCS_CHECK(a16.ensure(static_cast<size_t>(M) * K * 2));
    TCS_CHECK(w16.ensure(static_cast<size_t>(N) * K * 2));
    TCS_CHECK(launch_f32_to_bf16(A, a16.as<__nv_bfloat16>(), static_cast<size_t>(M) * K, st));
    TCS_CHECK(launch_f32_to_bf16(W, w16.as<__nv_bfloat16>(), static_cast<size_t>(N) * K, st));
    LinearTcPlan pl;
    TCS_CHECK(linear_tc_make_plan(&pl, a16.as<__nv_bfloat16>(), K, w16.as<__nv_bfloat16>(), K, M, N, K, bias, dst, N, flags, sms));
    TCS_CHECK(linear_tc_launch(pl, st));
  } else if (engine == TCS_ENGINE_SIMT) {
    TCS_CHECK(launch_linear_simt(A, K, W, K, M, N, K, bias, dst, N, flags, st));
  } else {
    return fail(TCS_ERR_BAD_ARGUMENT, "tcs_debug_linear: engine must be TCS_ENGINE_SIMT or TCS_ENGINE_TCGEN05");
  }
  if (bf16_out) {
    bf16_to_f32_kernel<<<static_cast<unsigned>((mn + 255) / 256), 256, 0, st>>>(o16.as<__nv_bfloat16>(), out, mn);
    TCS_CUDA(cudaGetLastError());
  }
  TCS_CUDA(cudaStreamSynchronize(st));
  return TCS_OK;
}

}  // extern "C"
