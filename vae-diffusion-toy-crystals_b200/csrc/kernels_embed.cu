// kernels_embed.cu — time / condition embeddings folded into a per-image bias of the first conv.
//
// Reference: timestep_embedding (sde_score_model.py:17-32), ConditionEmbedding.forward (:69-82),
// CondUNetTiny._make_maps (:227-241).  The reference materialises 16 spatially-constant feature
// maps and concatenates them to x; with circular padding a constant map contributes
// sum_taps(w) * value to every output pixel, so the 17->96 conv equals a 1->96 conv plus the
// per-(image, out-channel) bias computed here.
#include "kernels.cuh"

namespace tcs {

__device__ __forceinline__ float silu_acc(float v) { return v / (1.0f + expf(-v)); }

// y[o] = b[o] + sum_k w[o][k] * x[k]   (one thread per output)
__device__ __forceinline__ float dot_row(const float* __restrict__ w, const float* x, int K) {
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(__ldg(w + k), x[k], acc);
  return acc;
}

__global__ void __launch_bounds__(128) cond_embed_kernel(EmbedWeights w, const int64_t* __restrict__ y_cat,
                                                        const float* __restrict__ y_cont, int dup,
                                                        float* __restrict__ cvec) {
  __shared__ float yv[16];
  __shared__ float h0[128];
  __shared__ float z[256];
  __shared__ float ce[128];
  __shared__ float cmap[8];
  const int r = blockIdx.x, t = threadIdx.x;
  const int i = r / dup;
  const bool uncond = (dup == 2) && (r & 1);
  if (t < w.y_cont_dim) yv[t] = uncond ? 0.f : y_cont[static_cast<size_t>(i) * w.y_cont_dim + t];
  __syncthreads();
  if (t == 0) {
    const float s = sinf(yv[1]);
    yv[1] = s;
    yv[2] = cosf(s);  // the reference reads the already overwritten column (:76-78)
  }
  __syncthreads();
  long long cat = uncond ? w.n_types : y_cat[i];
  cat = cat < 0 ? 0 : (cat > w.n_types ? w.n_types : cat);
  h0[t] = silu_acc(dot_row(w.cm0_w + t * w.y_cont_dim, yv, w.y_cont_dim) + w.cm0_b[t]);
  __syncthreads();
  z[t] = silu_acc(w.cat_emb[cat * 128 + t]);
  z[128 + t] = silu_acc(dot_row(w.cm2_w + t * 128, h0, 128) + w.cm2_b[t]);
  __syncthreads();
  ce[t] = dot_row(w.co_w + t * 256, z, 256) + w.co_b[t];
  __syncthreads();
  if (t < 8) cmap[t] = dot_row(w.tc_w + t * 128, ce, 128) + w.tc_b[t];
  __syncthreads();
  if (t < 96) {
    float acc = 0.f;
    for (int j = 0; j < 8; ++j) acc = fmaf(w.wsum[t * 16 + 8 + j], cmap[j], acc);
    cvec[static_cast<size_t>(r) * 96 + t] = acc;
  }
}

__global__ void __launch_bounds__(128) time_embed_kernel(EmbedWeights w, const float* __restrict__ tvals,
                                                        float* __restrict__ tvec) {
  __shared__ float emb[128];
  __shared__ float h[128];
  __shared__ float te[128];
  __shared__ float tmap[8];
  const int r = blockIdx.x, t = threadIdx.x;
  const float tv = tvals[r];
  {
    const int i = t & 63;
    const float f = expf((-9.210340371976184f * static_cast<float>(i)) / 63.0f);
    const float a = (6.283185307179586f * tv) * f;
    emb[t] = t < 64 ? cosf(a) : sinf(a);
  }
  __syncthreads();
  h[t] = silu_acc(dot_row(w.tm0_w + t * 128, emb, 128) + w.tm0_b[t]);
  __syncthreads();
  te[t] = dot_row(w.tm2_w + t * 128, h, 128) + w.tm2_b[t];
  __syncthreads();
  if (t < 8) tmap[t] = dot_row(w.tt_w + t * 128, te, 128) + w.tt_b[t];
  __syncthreads();
  if (t < 96) {
    float acc = w.b0[t];
    for (int j = 0; j < 8; ++j) acc = fmaf(w.wsum[t * 16 + j], tmap[j], acc);
    tvec[static_cast<size_t>(r) * 96 + t] = acc;
  }
}

int launch_cond_embed(const EmbedWeights& w, const int64_t* y_cat, const float* y_cont, int n, int dup, float* cvec,
                      cudaStream_t st) {
  if (n <= 0) return TCS_OK;
  cond_embed_kernel<<<n * dup, 128, 0, st>>>(w, y_cat, y_cont, dup, cvec);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

int launch_time_embed(const EmbedWeights& w, const float* t, int nt, float* tvec, cudaStream_t st) {
  if (nt <= 0) return TCS_OK;
  time_embed_kernel<<<nt, 128, 0, st>>>(w, t, tvec);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

}  // namespace tcs
