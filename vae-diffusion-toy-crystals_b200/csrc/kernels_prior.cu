// kernels_prior.cu — CUDA-core kernels of the latent diffusion prior (FiLM residual MLP, DDIM eta = 0) and of the
// CondVAE decoder.  Reference: src/toycrystals/models/diffusion_prior.py (timestep_embedding :11-25, FiLMResBlock
// :39-54, DiffusionPriorFiLM.forward :108-127, ddim_sample :200-252) and src/toycrystals/models/vae.py (:36-43, 62-70).
// The dense layers of the FiLM blocks run on the tensor cores (linear_tc.cu) in bf16 mode; everything here is fp32.
#include "philox.cuh"
#include "prior.cuh"

namespace tcs {

__device__ __forceinline__ float silu_precise(float v) { return v / (1.0f + expf(-v)); }
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// fp32 GEMM, 64 x (16*TN) x 16 tiles, 4 x TN micro-tiles.  GATHER = the A rows are the 2x2-tap windows of a
// ConvTranspose2d(4, 2, 1) parity class (zero outside the image).
// ------------------------------------------------------------------------------------------------
struct GemmSimtArgs {
  const float* A; int lda;
  const float* W; int ldw;
  int M, N, K;
  const float* bias;
  void* out; int ldo;
  int flags;          // LinFlags; bit 8: ReLU (decoder)
  // GATHER geometry
  int Hi, Ci;
};
constexpr int GS_BM = 64, GS_BK = 16, GS_PAD = 4;
constexpr int LIN_RELU = 256;

template <int TN, bool GATHER>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmSimtArgs a) {
  constexpr int BN = 16 * TN;
  __shared__ __align__(16) float As[GS_BK][GS_BM + GS_PAD];
  __shared__ __align__(16) float Ws[GS_BK][BN + GS_PAD];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * GS_BM, n0 = blockIdx.y * BN;
  const int parity = GATHER ? blockIdx.z : 0;
  const int py = parity >> 1, px = parity & 1;
  // loader roles: A: thread -> (row t/4, k quad t%4); W: (row t/4 < BN, k quad)
  const int lr = t >> 2, lk = (t & 3) * 4;
  const int arow = m0 + lr;
  const float* aptr = nullptr;          // linear: row pointer
  const float* tap_ptr[4] = {nullptr, nullptr, nullptr, nullptr};
  if (GATHER) {
    if (arow < a.M) {
      const int Hi = a.Hi;
      const int b = arow / (Hi * Hi), rem = arow - b * Hi * Hi;
      const int i = rem / Hi, j = rem - i * Hi;
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) {
        const int ty = tp >> 1, tx = tp & 1;
        const int iy = i + (py == 0 ? (ty == 0 ? 0 : -1) : (ty == 0 ? 1 : 0));
        const int ix = j + (px == 0 ? (tx == 0 ? 0 : -1) : (tx == 0 ? 1 : 0));
        if (iy >= 0 && iy < Hi && ix >= 0 && ix < Hi)
          tap_ptr[tp] = a.A + (static_cast<size_t>(b) * Hi * Hi + static_cast<size_t>(iy) * Hi + ix) * a.Ci;
      }
    }
  } else if (arow < a.M) {
    aptr = a.A + static_cast<size_t>(arow) * a.lda;
  }
  const float* wbase = a.W + (GATHER ? static_cast<size_t>(parity) * a.N * a.K : 0);
  const float* wptr = (lr < BN && n0 + lr < a.N) ? wbase + static_cast<size_t>(n0 + lr) * a.ldw : nullptr;

  const int ty4 = (t >> 4) * 4, tx = (t & 15) * TN;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += GS_BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (GATHER) {
      const int kk = k0 + lk;
      const int tp = kk / a.Ci;
      const float* p = tp == 0 ? tap_ptr[0] : (tp == 1 ? tap_ptr[1] : (tp == 2 ? tap_ptr[2] : tap_ptr[3]));
      if (p) av = __ldg(reinterpret_cast<const float4*>(p + (kk - tp * a.Ci)));
    } else if (aptr) {
      av = __ldg(reinterpret_cast<const float4*>(aptr + k0 + lk));
    }
    if (wptr) wv = __ldg(reinterpret_cast<const float4*>(wptr + k0 + lk));
    __syncthreads();   // the previous slab has been consumed
    As[lk][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    if (lr < BN) { Ws[lk][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GS_BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty4]);
      float bv[TN];
      if (TN == 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Ws[k][tx]);
        bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
      } else {
        const float2 b2 = *reinterpret_cast<const float2*>(&Ws[k][tx]);
        bv[0] = b2.x; bv[1] = b2.y;
      }
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], bv[j], acc[i][j]);
    }
  }

  const int col = n0 + tx;
  if (col >= a.N) return;
  float bv[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) bv[j] = a.bias ? __ldg(a.bias + col + j) : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty4 + i;
    if (row >= a.M) continue;
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      v[j] = acc[i][j] + bv[j];
      if (a.flags & LIN_SILU) v[j] = silu_precise(v[j]);
      if (a.flags & LIN_RELU) v[j] = fmaxf(v[j], 0.f);
    }
    size_t orow;
    if (GATHER) {
      const int Hi = a.Hi, Ho = 2 * Hi;
      const int b = row / (Hi * Hi), rem = row - b * Hi * Hi;
      const int ii = rem / Hi, jj = rem - ii * Hi;
      orow = (static_cast<size_t>(b) * Ho + 2 * ii + py) * Ho + 2 * jj + px;
    } else {
      orow = row;
    }
    if (a.flags & LIN_OUT_F32) {
      float* o = static_cast<float*>(a.out) + orow * a.ldo + col;
      if (a.flags & LIN_ACCUM) {
#pragma unroll
        for (int j = 0; j < TN; ++j) v[j] += o[j];
      }
      if (TN == 4) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      else *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
    } else {
      __nv_bfloat16* o = static_cast<__nv_bfloat16*>(a.out) + orow * a.ldo + col;
      if (TN == 4) *reinterpret_cast<uint2*>(o) = make_uint2(pack2_bf16(v[0], v[1]), pack2_bf16(v[2], v[3]));
      else *reinterpret_cast<uint32_t*>(o) = pack2_bf16(v[0], v[1]);
    }
  }
}

int launch_linear_simt(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const float* bias, void* out,
                       int ldo, int flags, cudaStream_t st) {
  if (M < 1) return TCS_OK;
  if (N % 32 || K % GS_BK || (lda & 3) || (ldw & 3) || (ldo & 3))
    return fail(TCS_ERR_UNSUPPORTED, "linear_simt: needs N % 32 == 0, K % 16 == 0 and 16-byte row pitches");
  GemmSimtArgs a{};
  a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.M = M; a.N = N; a.K = K; a.bias = bias; a.out = out; a.ldo = ldo; a.flags = flags;
  if (N % 64 == 0) {
    gemm_simt_kernel<4, false><<<dim3((M + GS_BM - 1) / GS_BM, N / 64), 256, 0, st>>>(a);
  } else {
    gemm_simt_kernel<2, false><<<dim3((M + GS_BM - 1) / GS_BM, N / 32), 256, 0, st>>>(a);
  }
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

// ------------------------------------------------------------------------------------------------
// features
// ------------------------------------------------------------------------------------------------
__global__ void prior_ycat_kernel(PriorEmbedWeights w, const int64_t* __restrict__ y_cat, const float* __restrict__ y_cont,
                                  float* __restrict__ out) {
  extern __shared__ float sm[];   // yv[ycd] | h1[E]
  float* yv = sm;
  float* h1 = sm + 16;
  const int r = blockIdx.x, t = threadIdx.x, E = w.E;
  if (t < w.y_cont_dim) yv[t] = y_cont[static_cast<size_t>(r) * w.y_cont_dim + t];
  __syncthreads();
  long long cat = y_cat[r];
  cat = cat < 0 ? 0 : (cat >= w.n_types ? w.n_types - 1 : cat);
  float acc = w.cm0_b[t];
  for (int k = 0; k < w.y_cont_dim; ++k) acc = fmaf(__ldg(w.cm0_w + t * w.y_cont_dim + k), yv[k], acc);
  h1[t] = silu_precise(acc);
  __syncthreads();
  float o = w.cm2_b[t];
  for (int k = 0; k < E; ++k) o = fmaf(__ldg(w.cm2_w + t * E + k), h1[k], o);
  out[static_cast<size_t>(r) * 2 * E + t] = w.cat_emb[cat * E + t];
  out[static_cast<size_t>(r) * 2 * E + E + t] = o;
}
int launch_prior_ycat(const PriorEmbedWeights& w, const int64_t* y_cat, const float* y_cont, int n, float* out, cudaStream_t st) {
  if (n < 1) return TCS_OK;
  prior_ycat_kernel<<<n, w.E, (16 + w.E) * 4, st>>>(w, y_cat, y_cont, out);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

__global__ void prior_time_features_kernel(const void* __restrict__ t, int t_is_i32, const float* __restrict__ freqs, int rows,
                                           int dim, float* __restrict__ te) {
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * half) return;
  const int r = idx / half, i = idx - r * half;
  const float tv = t_is_i32 ? static_cast<float>(static_cast<const int*>(t)[r])
                            : static_cast<float>(static_cast<const long long*>(t)[r]);
  const float arg = tv * freqs[i];
  te[static_cast<size_t>(r) * dim + i] = sinf(arg);
  te[static_cast<size_t>(r) * dim + half + i] = cosf(arg);
  if ((dim & 1) && i == 0) te[static_cast<size_t>(r) * dim + dim - 1] = 0.f;
}
int launch_prior_time_features(const void* t, int t_is_i32, const float* freqs, int rows, int dim, float* te, cudaStream_t st) {
  if (rows < 1) return TCS_OK;
  const int total = rows * (dim / 2);
  prior_time_features_kernel<<<(total + 127) / 128, 128, 0, st>>>(t, t_is_i32, freqs, rows, dim, te);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm helpers: one warp per row, NV float4 per lane (W = 128 * NV); element c = i*128 + lane*4 + k
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void warp_layernorm(const float* __restrict__ row, int lane, const float* __restrict__ lw,
                                               const float* __restrict__ lb, float4 (&u)[NV]) {
  constexpr float invW = 1.0f / (128.0f * NV);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    u[i] = *reinterpret_cast<const float4*>(row + i * 128 + lane * 4);
    s += (u[i].x + u[i].y) + (u[i].z + u[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * invW;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float dx = u[i].x - mean, dy = u[i].y - mean, dz = u[i].z - mean, dw = u[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q * invW + LN_EPS);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(lw + i * 128 + lane * 4));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(lb + i * 128 + lane * 4));
    u[i].x = (u[i].x - mean) * rstd * w4.x + b4.x;
    u[i].y = (u[i].y - mean) * rstd * w4.y + b4.y;
    u[i].z = (u[i].z - mean) * rstd * w4.z + b4.z;
    u[i].w = (u[i].w - mean) * rstd * w4.w + b4.w;
  }
}

__device__ __forceinline__ float4 load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <int NV, typename TO>
__global__ void __launch_bounds__(256) ln_film_kernel(const float* __restrict__ h, int n, const float* __restrict__ ln_w,
                                                     const float* __restrict__ ln_b, const TO* __restrict__ film_row,
                                                     int film_ld, const float* __restrict__ film_step, int film_step_ld,
                                                     const int* __restrict__ step_ptr, int off, TO* __restrict__ out) {
  constexpr int W = 128 * NV;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  // the FiLM rows are independent of the LayerNorm result: issue their loads first so that both DRAM round trips overlap
  const TO* fr = film_row ? film_row + static_cast<size_t>(r) * film_ld + off : nullptr;
  const float* fs = film_step ? film_step + static_cast<size_t>(step_ptr ? *step_ptr : 0) * film_step_ld + off : nullptr;
  float4 g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    g[i] = fr ? load4(fr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    b[i] = fr ? load4(fr + W + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (fs) {
      const float4 g2 = __ldg(reinterpret_cast<const float4*>(fs + c));
      const float4 b2 = __ldg(reinterpret_cast<const float4*>(fs + W + c));
      g[i].x += g2.x; g[i].y += g2.y; g[i].z += g2.z; g[i].w += g2.w;
      b[i].x += b2.x; b[i].y += b2.y; b[i].z += b2.z; b[i].w += b2.w;
    }
  }
  float4 u[NV];
  warp_layernorm<NV>(h + static_cast<size_t>(r) * W, lane, ln_w, ln_b, u);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    const float o0 = u[i].x * (1.0f + g[i].x) + b[i].x, o1 = u[i].y * (1.0f + g[i].y) + b[i].y;
    const float o2 = u[i].z * (1.0f + g[i].z) + b[i].z, o3 = u[i].w * (1.0f + g[i].w) + b[i].w;
    if constexpr (sizeof(TO) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + static_cast<size_t>(r) * W + c) = make_float4(o0, o1, o2, o3);
    } else {
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + static_cast<size_t>(r) * W + c) =
          make_uint2(pack2_bf16(o0, o1), pack2_bf16(o2, o3));
    }
  }
}

template <typename TO>
int launch_ln_film(const float* h, int n, int W, const float* ln_w, const float* ln_b, const TO* film_row, int film_ld,
                   const float* film_step, int film_step_ld, const int* step_ptr, int off, TO* out, cudaStream_t st) {
  if (n < 1) return TCS_OK;
  const dim3 grid((n + 7) / 8);
#define TCS_LNF(NV) \
  ln_film_kernel<NV, TO><<<grid, 256, 0, st>>>(h, n, ln_w, ln_b, film_row, film_ld, film_step, film_step_ld, step_ptr, off, out)
  switch (W) {
    case 256: TCS_LNF(2); break;
    case 512: TCS_LNF(4); break;
    case 1024: TCS_LNF(8); break;
    case 2048: TCS_LNF(16); break;
    default: return fail(TCS_ERR_UNSUPPORTED, "ln_film: width must be 256, 512, 1024 or 2048");
  }
#undef TCS_LNF
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_ln_film<float>(const float*, int, int, const float*, const float*, const float*, int, const float*, int,
                                   const int*, int, float*, cudaStream_t);
template int launch_ln_film<__nv_bfloat16>(const float*, int, int, const float*, const float*, const __nv_bfloat16*, int,
                                           const float*, int, const int*, int, __nv_bfloat16*, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// tail of an evaluation: out_norm -> out_proj -> DDIM update -> in_proj of the next evaluation
// 16 rows per CTA (one warp per row), 512 threads
// ------------------------------------------------------------------------------------------------
constexpr int TAIL_ROWS = 16;

// Phases (one CTA = 16 rows, 16 warps):
//   1. warp w: LayerNorm(out_norm) of row w -> u_s[w][:] (shared memory, fp32)
//   2. warp w: out_proj outputs 2w, 2w+1 for all 16 rows (the two weight rows live in registers, so every CTA reads
//      out_proj.weight exactly once instead of once per row)
//   3. thread (r, j): DDIM update of z[r][j]
//   4. thread c: in_proj column c for the 16 rows (weights of the column in registers)
template <int NV>
__global__ void __launch_bounds__(512) prior_tail_kernel(const PriorTailArgs a) {
  constexpr int W = 128 * NV;
  extern __shared__ __align__(16) float u_s[];          // [TAIL_ROWS][W]
  __shared__ __align__(16) float zs[TAIL_ROWS][PRIOR_MAX_Z];
  __shared__ float eps_s[TAIL_ROWS][PRIOR_MAX_Z];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * TAIL_ROWS;
  const int zd = a.zd;
  const int step = a.step_ptr ? *a.step_ptr : 0;
  zs[warp][lane] = 0.f;   // columns >= zd and rows >= n stay zero
  __syncthreads();

  if (a.mode == TAIL_INIT) {
    const int r = row0 + warp;
    if (r < a.n && lane < zd) zs[warp][lane] = a.z[static_cast<size_t>(r) * zd + lane];
  } else {
    {   // phase 1
      const int r = row0 + warp;
      float4 u[NV];
      if (r < a.n) {
        warp_layernorm<NV>(a.h + static_cast<size_t>(r) * W, lane, a.on_w, a.on_b, u);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) u[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(u_s + warp * W + i * 128 + lane * 4) = u[i];
    }
    __syncthreads();
    // phase 2: every row of u_s is read once per warp and used for both of the warp's outputs
    {
      const int o0 = warp * 2;
      float4 wa[NV], wb[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        wa[i] = o0 < zd ? __ldg(reinterpret_cast<const float4*>(a.op_w + static_cast<size_t>(o0) * W + i * 128 + lane * 4))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        wb[i] = o0 + 1 < zd ? __ldg(reinterpret_cast<const float4*>(a.op_w + static_cast<size_t>(o0 + 1) * W + i * 128 + lane * 4))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float mine_a = 0.f, mine_b = 0.f;   // lane r keeps the results of row r
#pragma unroll 2
      for (int r = 0; r < TAIL_ROWS; ++r) {
        float acc_a = 0.f, acc_b = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 x = *reinterpret_cast<const float4*>(u_s + r * W + i * 128 + lane * 4);
          acc_a = fmaf(x.x, wa[i].x, acc_a); acc_a = fmaf(x.y, wa[i].y, acc_a);
          acc_a = fmaf(x.z, wa[i].z, acc_a); acc_a = fmaf(x.w, wa[i].w, acc_a);
          acc_b = fmaf(x.x, wb[i].x, acc_b); acc_b = fmaf(x.y, wb[i].y, acc_b);
          acc_b = fmaf(x.z, wb[i].z, acc_b); acc_b = fmaf(x.w, wb[i].w, acc_b);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          acc_a += __shfl_xor_sync(0xffffffffu, acc_a, off);
          acc_b += __shfl_xor_sync(0xffffffffu, acc_b, off);
        }
        if (lane == r) { mine_a = acc_a; mine_b = acc_b; }
      }
      if (lane < TAIL_ROWS) {
        if (o0 < zd) eps_s[lane][o0] = mine_a + __ldg(a.op_b + o0);
        if (o0 + 1 < zd) eps_s[lane][o0 + 1] = mine_b + __ldg(a.op_b + o0 + 1);
      }
    }
    __syncthreads();
    // phase 3
    const int r = row0 + warp;
    if (r < a.n && lane < zd) {
      const float eps = eps_s[warp][lane];
      const size_t zi = static_cast<size_t>(r) * zd + lane;
      if (a.mode == TAIL_EPS) {
        a.eps_out[zi] = eps;
      } else {
        const DdimCoef c = a.coef[step];
        const float z = a.z[zi];
        const size_t ti = (static_cast<size_t>(step) * a.trace_n + r) * zd + lane;
        if (a.trace_eps) a.trace_eps[ti] = eps;
        if (a.trace_z) a.trace_z[ti] = z;
        // z0_pred = (z - sqrt(1-abar) eps) / (sqrt(abar) + 1e-8)            (:233-238)
        const float z0 = __fdiv_rn(__fsub_rn(z, __fmul_rn(c.s1m_t, eps)), c.sa_t_eps);
        float zn = z0;
        if (!c.last) zn = __fadd_rn(__fmul_rn(c.sa_prev, z0), __fmul_rn(c.s1m_prev, eps));   // (:250)
        if (c.last) a.z_out[zi] = zn;
        else a.z[zi] = zn;
        zs[warp][lane] = zn;
      }
    }
    if (a.mode == TAIL_EPS) return;
    if (a.coef[step].last) return;   // uniform over the grid
  }
  __syncthreads();
  // phase 4: in_proj for the 16 rows of this CTA.  in_proj.weight [W][zd] is staged through the (now free) row buffer
  // in coalesced 16-byte pieces, XOR-swizzled so that thread c can read "its" row with conflict-free 128-bit loads;
  // W/2 columns per pass (zd = 32; smaller zd just leaves the tail of every row unused)
  constexpr int CPP = W / 2;                          // columns per pass: CPP * 32 floats = the row buffer
  float4* w_s = reinterpret_cast<float4*>(u_s);
  const int zq = zd >> 2;                             // float4 per weight row
  for (int c0 = 0; c0 < W; c0 += CPP) {
    __syncthreads();                                  // the previous contents of the buffer are dead
    for (int f = threadIdx.x; f < CPP * 8; f += 512) {
      const int cl = f >> 3, j4 = f & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j4 < zq) v = __ldg(reinterpret_cast<const float4*>(a.ip_w + static_cast<size_t>(c0 + cl) * zd) + j4);
      w_s[cl * 8 + (j4 ^ (cl & 7))] = v;
    }
    __syncthreads();
    for (int cl = threadIdx.x; cl < CPP; cl += 512) {
      float4 w[8];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) w[j4] = w_s[cl * 8 + (j4 ^ (cl & 7))];
      const int c = c0 + cl;
      const float b = __ldg(a.ip_b + c);
      for (int rr = 0; rr < TAIL_ROWS; ++rr) {
        if (row0 + rr >= a.n) break;
        float acc = b;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 zv = *reinterpret_cast<const float4*>(&zs[rr][j4 * 4]);   // zero beyond zd
          acc = fmaf(w[j4].x, zv.x, acc); acc = fmaf(w[j4].y, zv.y, acc);
          acc = fmaf(w[j4].z, zv.z, acc); acc = fmaf(w[j4].w, zv.w, acc);
        }
        a.h[static_cast<size_t>(row0 + rr) * W + c] = acc;
      }
    }
  }
}

template <int NV>
static int launch_tail_t(const PriorTailArgs& a, cudaStream_t st) {
  auto kern = prior_tail_kernel<NV>;
  const size_t smem = static_cast<size_t>(TAIL_ROWS) * 128 * NV * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    TCS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_done = true;
  }
  kern<<<(a.n + TAIL_ROWS - 1) / TAIL_ROWS, 512, smem, st>>>(a);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

int launch_prior_tail(const PriorTailArgs& a, cudaStream_t st) {
  if (a.n < 1) return TCS_OK;
  if (a.zd > PRIOR_MAX_Z) return fail(TCS_ERR_UNSUPPORTED, "prior_tail: z_dim must be <= 32");
  switch (a.W) {
    case 256: return launch_tail_t<2>(a, st);
    case 512: return launch_tail_t<4>(a, st);
    case 1024: return launch_tail_t<8>(a, st);
    case 2048: return launch_tail_t<16>(a, st);
    default: return fail(TCS_ERR_UNSUPPORTED, "prior_tail: width must be 256, 512, 1024 or 2048");
  }
}

__global__ void __launch_bounds__(256) prior_philox_kernel(float4* __restrict__ z, long long nvec, int groups,
                                                          unsigned long long seed, unsigned long long gidx0) {
  const long long v = blockIdx.x * 256LL + threadIdx.x;
  if (v >= nvec) return;
  const long long s = v / groups;
  z[v] = philox_normal4(seed, gidx0 + static_cast<unsigned long long>(s), 0u, static_cast<uint32_t>(v - s * groups));
}
int launch_prior_philox(float* z, int n, int zd, unsigned long long seed, unsigned long long gidx0, cudaStream_t st) {
  if (n < 1) return TCS_OK;
  if (zd % 4) return fail(TCS_ERR_UNSUPPORTED, "prior_philox: z_dim must be a multiple of 4");
  const long long nvec = static_cast<long long>(n) * (zd / 4);
  prior_philox_kernel<<<static_cast<unsigned>((nvec + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(z), nvec, zd / 4,
                                                                              seed, gidx0);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                         size_t count) {
  const size_t i = (blockIdx.x * 256ULL + threadIdx.x) * 4;
  if (i + 3 < count) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack2_bf16(v.x, v.y), pack2_bf16(v.z, v.w));
  } else {
    for (size_t k = i; k < count; ++k) dst[k] = __float2bfloat16(src[k]);
  }
}
int launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, size_t count, cudaStream_t st) {
  if (!count) return TCS_OK;
  const size_t blocks = (count / 4 + 256) / 256;
  f32_to_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, dst, count);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

// ------------------------------------------------------------------------------------------------
// CondVAE decoder
// ------------------------------------------------------------------------------------------------
constexpr int DFC_IMGS = 8;
template <typename TO>
__global__ void __launch_bounds__(256) vae_dec_fc_kernel(const float* __restrict__ z, const int64_t* __restrict__ y_cat,
                                                        const float* __restrict__ y_cont, const float* __restrict__ z_mean,
                                                        const float* __restrict__ z_std, const float* __restrict__ w,
                                                        const float* __restrict__ b, int n, int zd, int n_types, int ycd,
                                                        TO* __restrict__ h0) {
  __shared__ float v[DFC_IMGS][64];   // [z' | y_cont] per image
  __shared__ int cat[DFC_IMGS];
  const int img0 = blockIdx.x * DFC_IMGS, t = threadIdx.x;
  const int in_dim = zd + n_types + ycd;
  for (int idx = t; idx < DFC_IMGS * (zd + ycd); idx += 256) {
    const int im = idx / (zd + ycd), k = idx - im * (zd + ycd);
    float val = 0.f;
    if (img0 + im < n) {
      if (k < zd) {
        val = z[static_cast<size_t>(img0 + im) * zd + k];
        if (z_std) val = __fadd_rn(__fmul_rn(val, z_std[k]), z_mean[k]);   // z = z_norm * z_std + z_mean
      } else {
        val = y_cont[static_cast<size_t>(img0 + im) * ycd + (k - zd)];
      }
    }
    v[im][k] = val;
  }
  if (t < DFC_IMGS) {
    long long c = img0 + t < n ? y_cat[img0 + t] : 0;
    cat[t] = static_cast<int>(c < 0 ? 0 : (c >= n_types ? n_types - 1 : c));
  }
  __syncthreads();
  const int c = t;   // output channel; output index o = c*16 + p (view(-1, 256, 4, 4))
  for (int p = 0; p < 16; ++p) {
    const float* wr = w + static_cast<size_t>(c * 16 + p) * in_dim;
    float acc[DFC_IMGS];
    const float bias = __ldg(b + c * 16 + p);
#pragma unroll
    for (int im = 0; im < DFC_IMGS; ++im) acc[im] = bias;
    for (int k = 0; k < zd; ++k) {
      const float wk = __ldg(wr + k);
#pragma unroll
      for (int im = 0; im < DFC_IMGS; ++im) acc[im] = fmaf(wk, v[im][k], acc[im]);
    }
    for (int k = 0; k < ycd; ++k) {
      const float wk = __ldg(wr + zd + n_types + k);
#pragma unroll
      for (int im = 0; im < DFC_IMGS; ++im) acc[im] = fmaf(wk, v[im][zd + k], acc[im]);
    }
#pragma unroll
    for (int im = 0; im < DFC_IMGS; ++im) {
      if (img0 + im >= n) break;
      const float val = acc[im] + __ldg(wr + zd + cat[im]);
      if constexpr (sizeof(TO) == 4) h0[(static_cast<size_t>(img0 + im) * 16 + p) * 256 + c] = val;
      else h0[(static_cast<size_t>(img0 + im) * 16 + p) * 256 + c] = __float2bfloat16(val);
    }
  }
}
int launch_vae_dec_fc(const float* z, const int64_t* y_cat, const float* y_cont, const float* z_mean, const float* z_std,
                      const float* w, const float* b, int n, int zd, int n_types, int ycd, void* h0, int out_bf16,
                      cudaStream_t st) {
  if (n < 1) return TCS_OK;
  if (zd + ycd > 64) return fail(TCS_ERR_UNSUPPORTED, "vae_dec_fc: z_dim + y_cont_dim must be <= 64");
  const dim3 grid((n + DFC_IMGS - 1) / DFC_IMGS);
  if (out_bf16)
    vae_dec_fc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(z, y_cat, y_cont, z_mean, z_std, w, b, n, zd, n_types, ycd,
                                                          static_cast<__nv_bfloat16*>(h0));
  else
    vae_dec_fc_kernel<float><<<grid, 256, 0, st>>>(z, y_cat, y_cont, z_mean, z_std, w, b, n, zd, n_types, ycd, static_cast<float*>(h0));
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

// tap (ty, tx) of output parity (py, px): ky = py ? (ty ? 2 : 0) : (ty ? 3 : 1), likewise kx
static inline int convt_k(int parity_bit, int tap_bit) { return parity_bit ? (tap_bit ? 2 : 0) : (tap_bit ? 3 : 1); }

void vae_convt_pack_weights(const float* w, int Ci, int Co, float* out) {
  // [parity][co][tap*Ci + ci]  <-  w[ci][co][ky][kx]
  for (int par = 0; par < 4; ++par)
    for (int co = 0; co < Co; ++co)
      for (int tp = 0; tp < 4; ++tp) {
        const int ky = convt_k(par >> 1, tp >> 1), kx = convt_k(par & 1, tp & 1);
        for (int ci = 0; ci < Ci; ++ci)
          out[(static_cast<size_t>(par) * Co + co) * 4 * Ci + tp * Ci + ci] = w[((static_cast<size_t>(ci) * Co + co) * 4 + ky) * 4 + kx];
      }
}

int launch_vae_convt(const float* in, const float* wpacked, const float* bias, int n, int Hi, int Ci, int Co, float* out,
                     cudaStream_t st) {
  if (n < 1) return TCS_OK;
  if (Ci % 16 || Co % 32) return fail(TCS_ERR_UNSUPPORTED, "vae_convt: needs C_in % 16 == 0 and C_out % 32 == 0");
  GemmSimtArgs a{};
  a.A = in; a.lda = 0; a.W = wpacked; a.ldw = 4 * Ci; a.M = n * Hi * Hi; a.N = Co; a.K = 4 * Ci; a.bias = bias; a.out = out;
  a.ldo = Co; a.flags = LIN_OUT_F32 | LIN_RELU; a.Hi = Hi; a.Ci = Ci;
  const int mt = (a.M + GS_BM - 1) / GS_BM;
  if (Co % 64 == 0) gemm_simt_kernel<4, true><<<dim3(mt, Co / 64, 4), 256, 0, st>>>(a);
  else gemm_simt_kernel<2, true><<<dim3(mt, Co / 32, 4), 256, 0, st>>>(a);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

template <typename TI>
__global__ void __launch_bounds__(256) vae_convt_out_kernel(const TI* __restrict__ in, const float* __restrict__ wpacked,
                                                           float bias, long long total, float* __restrict__ x) {
  __shared__ __align__(16) float ws[4 * 4 * 32];
  for (int i = threadIdx.x; i < 512; i += 256) ws[i] = wpacked[i];
  __syncthreads();
  const long long idx = blockIdx.x * 256LL + threadIdx.x;
  if (idx >= total) return;
  const int ox = static_cast<int>(idx & 63), oy = static_cast<int>((idx >> 6) & 63);
  const long long b = idx >> 12;
  const int py = oy & 1, px = ox & 1, i = oy >> 1, j = ox >> 1;
  float acc = bias;
#pragma unroll
  for (int tp = 0; tp < 4; ++tp) {
    const int ty = tp >> 1, tx = tp & 1;
    const int iy = i + (py == 0 ? (ty == 0 ? 0 : -1) : (ty == 0 ? 1 : 0));
    const int ix = j + (px == 0 ? (tx == 0 ? 0 : -1) : (tx == 0 ? 1 : 0));
    if (iy < 0 || iy >= 32 || ix < 0 || ix >= 32) continue;
    const float4* wv = reinterpret_cast<const float4*>(ws + ((py * 2 + px) * 4 + tp) * 32);
    if constexpr (sizeof(TI) == 4) {
      const float4* p = reinterpret_cast<const float4*>(in + ((b * 32 + iy) * 32 + ix) * 32);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 a4 = __ldg(p + k), w4 = wv[k];
        acc = fmaf(a4.x, w4.x, acc); acc = fmaf(a4.y, w4.y, acc); acc = fmaf(a4.z, w4.z, acc); acc = fmaf(a4.w, w4.w, acc);
      }
    } else {
      const uint4* p = reinterpret_cast<const uint4*>(in + ((b * 32 + iy) * 32 + ix) * 32);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 a8 = __ldg(p + k);
        const uint32_t aw[4] = {a8.x, a8.y, a8.z, a8.w};
        const float4 w0 = wv[2 * k], w1 = wv[2 * k + 1];
        const float wf[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[e]));
          acc = fmaf(f.x, wf[2 * e], acc); acc = fmaf(f.y, wf[2 * e + 1], acc);
        }
      }
    }
  }
  x[idx] = 1.0f / (1.0f + expf(-acc));
}
int launch_vae_convt_out(const void* in, int in_bf16, const float* wpacked, float bias, int n, float* x, cudaStream_t st) {
  if (n < 1) return TCS_OK;
  const long long total = static_cast<long long>(n) * 4096;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (in_bf16) vae_convt_out_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), wpacked, bias, total, x);
  else vae_convt_out_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in), wpacked, bias, total, x);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

}  // namespace tcs
