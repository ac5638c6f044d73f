// kernels_simt.cu — CUDA-core kernels of the score network (everything that is not a
// tensor-core GEMM): first 1->96 conv, GroupNorm(+SiLU), bilinear x2, attention, 96->1 out conv
// with the CFG combine, and a generic fp32-accumulate SIMT convolution that is the whole conv
// engine of the fp32 mode (and the on-GPU cross-check of the tcgen05 engine in bf16 mode).
//
// Reference: CondUNetTiny.forward, src/toycrystals/models/sde_score_model.py:243-266 and the
// modules it calls (_ConvBlock :97-111, SelfAttention2d :114-167, nn.Upsample :217,221).
#include "kernels.cuh"

#include <cstdlib>

namespace tcs {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
template <typename T> struct Vec8;  // 8 consecutive channels
template <> struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = a; *reinterpret_cast<float4*>(p + 4) = b; }
  __device__ __forceinline__ void get(float* f) const { f[0]=a.x; f[1]=a.y; f[2]=a.z; f[3]=a.w; f[4]=b.x; f[5]=b.y; f[6]=b.z; f[7]=b.w; }
  __device__ __forceinline__ void set(const float* f) { a = make_float4(f[0],f[1],f[2],f[3]); b = make_float4(f[4],f[5],f[6],f[7]); }
};
template <> struct Vec8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = u; }
  __device__ __forceinline__ void get(float* f) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2*i] = t.x; f[2*i+1] = t.y; }
  }
  __device__ __forceinline__ void set(const float* f) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2*i], f[2*i+1]);
  }
};

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// padded-tensor halo bookkeeping: image (y,x) lives at padded (y+1,x+1); rows/cols on the border
// are duplicated on the opposite halo.  wrap offset in padded rows/cols (0 = interior only).
__device__ __forceinline__ int halo_wrap(int v, int n) { return v == 0 ? n : (v == n - 1 ? -n : 0); }

// ------------------------------------------------------------------------------------------
// first conv: x[n,64,64] (*) w9[96][9], circular, + tvec[row] + cvec[b]   -> raw fp32 + GN partials
// block = 2 image rows (128 px) of one sample, 192 threads: oc = t%96, pixel parity = t/96
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192) first_conv_kernel(const float* __restrict__ x, const float* __restrict__ w9,
                                                        const float* __restrict__ tvec, int tvec_stride,
                                                        const int* __restrict__ step_ptr, int trow_off,
                                                        const float* __restrict__ cvec, int dup,
                                                        float* __restrict__ raw, float* __restrict__ partials) {
  __shared__ float xs[4][IMG + 2];
  __shared__ float red[2][96][2][2];  // [half][oc][u][sum,sumsq]
  const int i = blockIdx.x >> 5, rp = blockIdx.x & 31, y0 = rp * 2;
  const int t = threadIdx.x, oc = t % 96, half = t / 96;
  for (int e = t; e < 4 * (IMG + 2); e += 192) {
    const int r = e / (IMG + 2), c = e % (IMG + 2);
    const int yy = (y0 - 1 + r + IMG) & (IMG - 1), xx = (c - 1 + IMG) & (IMG - 1);
    xs[r][c] = x[(static_cast<size_t>(i) * IMG + yy) * IMG + xx];
  }
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = w9[oc * 9 + k];
  const int trow = (step_ptr ? *step_ptr : 0) + trow_off + i * tvec_stride;
  const float tb = tvec[static_cast<size_t>(trow) * 96 + oc];
  float bias[2], s[2] = {0.f, 0.f}, q[2] = {0.f, 0.f};
  for (int u = 0; u < dup; ++u) bias[u] = tb + cvec[(static_cast<size_t>(i) * dup + u) * 96 + oc];
  __syncthreads();
  for (int p = half; p < 128; p += 2) {
    const int r = p >> 6, c = p & 63;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) acc = fmaf(w[ky * 3 + kx], xs[r + ky][c + kx], acc);
    for (int u = 0; u < dup; ++u) {
      const float v = acc + bias[u];
      raw[((static_cast<size_t>(i) * dup + u) * IMG_PIX + (y0 + r) * IMG + c) * 96 + oc] = v;
      s[u] += v;
      q[u] += v * v;
    }
  }
  for (int u = 0; u < dup; ++u) { red[half][oc][u][0] = s[u]; red[half][oc][u][1] = q[u]; }
  __syncthreads();
  if (t < 8 * dup) {
    const int g = t % 8, u = t / 8;
    float ss = 0.f, qq = 0.f;
    for (int h = 0; h < 2; ++h)
      for (int c = 0; c < 12; ++c) { ss += red[h][g * 12 + c][u][0]; qq += red[h][g * 12 + c][u][1]; }
    float* dst = partials + ((static_cast<size_t>(i) * dup + u) * FIRST_CONV_SLOTS + rp) * 16 + 2 * g;
    dst[0] = ss;
    dst[1] = qq;
  }
}

int launch_first_conv(const float* x, const float* w9, const float* tvec, int tvec_stride, const int* step_ptr,
                      int trow_off, const float* cvec, int n, int dup, float* raw, float* partials,
                      cudaStream_t st) {
  if (n <= 0) return TCS_OK;
  first_conv_kernel<<<n * 32, 192, 0, st>>>(x, w9, tvec, tvec_stride, step_ptr, trow_off, cvec, dup, raw, partials);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

template <bool FAST> __device__ __forceinline__ float silu_f(float v);

// ------------------------------------------------------------------------------------------
// first conv + GroupNorm + SiLU in ONE kernel (down1.net.0 + down1.net.1 + SiLU): one block per
// sample; the 1->96 conv costs 9 MACs per output, so it is simply evaluated twice (pass 1: group
// statistics, pass 2: normalise + write the padded activation) instead of round-tripping fp32.
// thread = (output-channel pair, pixel phase): 48 x 8 = 384 threads.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 pk_fma2v(float2 a, float2 b, float2 c) {   // packed fp32 FMA (FFMA2)
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
template <typename T> __device__ __forceinline__ void store_pair(T* p, float a, float b);
template <> __device__ __forceinline__ void store_pair<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <> __device__ __forceinline__ void store_pair<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// Statistics without a first pass: with circular padding  sum_p conv_c(p) = (sum_k w_ck) * sum_p x_p  and
// sum_p conv_c(p)^2 = sum_{k,l} w_ck w_cl R(k-l)  with  R(d) = sum_p x_p x_{p+d}  (circular autocorrelation at
// the 25 lags |dy|,|dx| <= 2), both exact identities; R and the quadratic forms are evaluated in fp64.
// Each sample is split over FC_SPLIT blocks (row bands); every block recomputes the (cheap) statistics.
constexpr int FC_SPLIT = 2;
template <typename T>
__global__ void __launch_bounds__(384, 2) first_conv_gn_kernel(const float* __restrict__ x, const float* __restrict__ w9,
                                                           const float* __restrict__ tvec, int tvec_stride,
                                                           const int* __restrict__ step_ptr, int trow_off,
                                                           const float* __restrict__ cvec, int dup,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           T* __restrict__ out) {
  constexpr bool FAST = sizeof(T) == 2;
  constexpr int P = IMG + 4;                 // halo of 2 for the autocorrelation lags
  __shared__ float2 xs[P][P];                // {x, x}: the packed-FMA main loop wants the input value in both halves
  __shared__ double racc[12][14];            // per-warp partial: 13 lags + sum x
  __shared__ double R[5][5];
  __shared__ double S1;
  __shared__ double csum[96][2][2];          // per channel, per branch: sum v, sum v^2
  __shared__ float s_mean[2][GN_GROUPS], s_rstd[2][GN_GROUPS];
  const int i = blockIdx.x / FC_SPLIT, band = blockIdx.x % FC_SPLIT;
  const int t = threadIdx.x, op = t % 48, pg = t / 48, oc = 2 * op;
  const int warp = t >> 5, lane = t & 31;
  for (int e = t; e < P * P; e += 384) {
    const int r = e / P, c = e - r * P;
    const float xval = x[(static_cast<size_t>(i) * IMG + ((r - 2 + IMG) & (IMG - 1))) * IMG + ((c - 2 + IMG) & (IMG - 1))];
    xs[r][c] = make_float2(xval, xval);
  }
  __syncthreads();
  {  // R(dy,dx) for the 13 lags with (dy > 0) or (dy == 0 and dx >= 0); R(-d) = R(d)
    double acc[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) acc[k] = 0.0;
    for (int p = t; p < IMG_PIX; p += 384) {
      const int yy = (p >> 6) + 2, xx = (p & 63) + 2;
      const double x0 = xs[yy][xx].x;
      acc[13] += x0;
#pragma unroll
      for (int k = 0; k < 13; ++k) {
        const int dy = (k + 2) / 5, dx = (k + 2) % 5 - 2;   // k=0..2 -> dy=0,dx=0..2 ; then dy=1,2 with dx=-2..2
        acc[k] += x0 * static_cast<double>(xs[yy + dy][xx + dx].x);
      }
    }
#pragma unroll
    for (int k = 0; k < 14; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      if (lane == 0) racc[warp][k] = acc[k];
    }
  }
  __syncthreads();
  if (t < 14) {
    double v = 0.0;
    for (int w = 0; w < 12; ++w) v += racc[w][t];
    if (t == 13) S1 = v;
    else {
      const int dy = (t + 2) / 5, dx = (t + 2) % 5 - 2;
      R[2 + dy][2 + dx] = v;
      R[2 - dy][2 - dx] = v;
    }
  }
  const int trow = (step_ptr ? *step_ptr : 0) + trow_off + i * tvec_stride;
  __syncthreads();
  if (t < 96) {
    double wk[9], A = 0.0, Q = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) { wk[k] = w9[t * 9 + k]; A += wk[k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int l = 0; l < 9; ++l) Q += wk[k] * wk[l] * R[2 + k / 3 - l / 3][2 + k % 3 - l % 3];
    for (int u = 0; u < dup; ++u) {
      const double bb = static_cast<double>(tvec[static_cast<size_t>(trow) * 96 + t] + cvec[(static_cast<size_t>(i) * dup + u) * 96 + t]);
      csum[t][u][0] = A * S1 + IMG_PIX * bb;
      csum[t][u][1] = Q + 2.0 * bb * A * S1 + IMG_PIX * bb * bb;
    }
  }
  __syncthreads();
  if (t < GN_GROUPS * dup) {
    const int g = t & 7, u = t >> 3;
    double sd = 0.0, qd = 0.0;
    for (int j = 0; j < 12; ++j) { sd += csum[g * 12 + j][u][0]; qd += csum[g * 12 + j][u][1]; }
    const double cnt = static_cast<double>(IMG_PIX) * 12.0;
    const double mean = sd / cnt;
    double var = qd / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mean[u][g] = static_cast<float>(mean);
    s_rstd[u][g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(GN_EPS)));
  }
  float w0[9], w1[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { w0[k] = w9[oc * 9 + k]; w1[k] = w9[(oc + 1) * 9 + k]; }
  float b0[2], b1[2];
  for (int u = 0; u < 2; ++u) {
    const int uu = u < dup ? u : 0;
    b0[u] = tvec[static_cast<size_t>(trow) * 96 + oc] + cvec[(static_cast<size_t>(i) * dup + uu) * 96 + oc];
    b1[u] = tvec[static_cast<size_t>(trow) * 96 + oc + 1] + cvec[(static_cast<size_t>(i) * dup + uu) * 96 + oc + 1];
  }
  __syncthreads();
  const int g = op / 6;
  // y = conv * sc + sh with sc = rstd * gamma, sh = (bias - mean) * sc + beta   (per branch u, per channel)
  float sc0[2], sh0[2], sc1[2], sh1[2];
  {
    const float ga0 = gamma[oc], ga1 = gamma[oc + 1], be0 = beta[oc], be1 = beta[oc + 1];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      sc0[u] = s_rstd[u][g] * ga0; sh0[u] = (b0[u] - s_mean[u][g]) * sc0[u] + be0;
      sc1[u] = s_rstd[u][g] * ga1; sh1[u] = (b1[u] - s_mean[u][g]) * sc1[u] + be1;
    }
  }
  constexpr int PO = IMG + 2;
  const int p_begin = band * (IMG_PIX / FC_SPLIT), p_end = p_begin + IMG_PIX / FC_SPLIT;
  // runs of 4 horizontally adjacent pixels: 18 shared-memory loads feed 72 FMAs (the one-pixel version was LDS-bound)
  (void)p_end;
  if constexpr (FAST) {
    // bf16 output: packed fp32 math (FFMA2).  The two channels of a thread are the two halves of every operand: weights
    // {w0, w1}, input {x, x} (stored duplicated), accumulators {c0, c1}; 0.5 of SiLU(y) = h + h tanh(h), h = y / 2, is
    // folded into the affine.  (The scalar version was issue-bound: 77 % issue-active for 160 MB of output.)
    float2 w2[9], sc2[2], sh2[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) w2[k] = make_float2(w0[k], w1[k]);
#pragma unroll
    for (int u = 0; u < 2; ++u) { sc2[u] = make_float2(0.5f * sc0[u], 0.5f * sc1[u]); sh2[u] = make_float2(0.5f * sh0[u], 0.5f * sh1[u]); }
    for (int qd = pg; qd < IMG_PIX / FC_SPLIT / 4; qd += 8) {
      const int p = p_begin + qd * 4;
      const int yy = p >> 6, xx = p & 63;
      float2 c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float2 xv[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) xv[k] = xs[yy + 1 + ky][xx + 1 + k];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int j = 0; j < 4; ++j) c[j] = pk_fma2v(w2[ky * 3 + kx], xv[j + kx], c[j]);
      }
      // addresses: one 64-bit base per run, then constant strides (the scalar version spent a third of its issue slots
      // on predicated 64-bit address arithmetic for the halo copies)
      T* const o_run = out + ((static_cast<size_t>(i) * dup * PO + yy + 1) * PO + xx + 1) * 96 + oc;
      constexpr long long IMG_STRIDE = static_cast<long long>(PO) * PO * 96, COL_WRAP = static_cast<long long>(IMG) * 96;
      const long long wyo = static_cast<long long>(halo_wrap(yy, IMG)) * PO * 96;   // 0 = not a border row
      const bool left = xx == 0, right = xx == IMG - 4;                           // runs are 4-aligned
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u >= dup) break;
          const float2 hh = pk_fma2v(c[j], sc2[u], sh2[u]);
          float t0, t1;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(hh.x));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(hh.y));
          const float2 yv = pk_fma2v(hh, make_float2(t0, t1), hh);
          T* const o = o_run + j * 96 + u * IMG_STRIDE;
          store_pair<T>(o, yv.x, yv.y);
          if (wyo) store_pair<T>(o + wyo, yv.x, yv.y);
          if (j == 0 && left) {
            store_pair<T>(o + COL_WRAP, yv.x, yv.y);
            if (wyo) store_pair<T>(o + wyo + COL_WRAP, yv.x, yv.y);
          }
          if (j == 3 && right) {
            store_pair<T>(o - COL_WRAP, yv.x, yv.y);
            if (wyo) store_pair<T>(o + wyo - COL_WRAP, yv.x, yv.y);
          }
        }
      }
    }
    return;
  }
  for (int qd = pg; qd < IMG_PIX / FC_SPLIT / 4; qd += 8) {
    const int p = p_begin + qd * 4;
    const int yy = p >> 6, xx = p & 63;
    float xv[3][6];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int k = 0; k < 6; ++k) xv[ky][k] = xs[yy + 1 + ky][xx + 1 + k].x;
    float c[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          c[j][0] = fmaf(w0[ky * 3 + kx], xv[ky][j + kx], c[j][0]);
          c[j][1] = fmaf(w1[ky * 3 + kx], xv[ky][j + kx], c[j][1]);
        }
    const int wy = halo_wrap(yy, IMG);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int wx = halo_wrap(xx + j, IMG);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u >= dup) break;
        const float y0 = silu_f<FAST>(fmaf(c[j][0], sc0[u], sh0[u]));
        const float y1 = silu_f<FAST>(fmaf(c[j][1], sc1[u], sh1[u]));
        const size_t base = ((static_cast<size_t>(i) * dup + u) * PO + yy + 1) * PO + xx + j + 1;
        store_pair<T>(out + base * 96 + oc, y0, y1);
        if (wy) store_pair<T>(out + (base + static_cast<long long>(wy) * PO) * 96 + oc, y0, y1);
        if (wx) store_pair<T>(out + (base + wx) * 96 + oc, y0, y1);
        if (wy && wx) store_pair<T>(out + (base + static_cast<long long>(wy) * PO + wx) * 96 + oc, y0, y1);
      }
    }
  }
}

template <typename T>
int launch_first_conv_gn(const float* x, const float* w9, const float* tvec, int tvec_stride, const int* step_ptr,
                         int trow_off, const float* cvec, int n, int dup, const float* gamma, const float* beta, T* out,
                         cudaStream_t st) {
  if (n <= 0) return TCS_OK;
  first_conv_gn_kernel<T><<<n * FC_SPLIT, 384, 0, st>>>(x, w9, tvec, tvec_stride, step_ptr, trow_off, cvec, dup, gamma,
                                                        beta, out);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_first_conv_gn<float>(const float*, const float*, const float*, int, const int*, int, const float*, int, int, const float*, const float*, float*, cudaStream_t);
template int launch_first_conv_gn<__nv_bfloat16>(const float*, const float*, const float*, int, const int*, int, const float*, int, int, const float*, const float*, __nv_bfloat16*, cudaStream_t);

// ------------------------------------------------------------------------------------------
// GroupNorm statistics from partial sums (fp64 combine, fixed order -> deterministic)
// one block per image: stats[b][g] = (mean, rstd)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ partials, int slots, double count,
                                                         float2* __restrict__ stats) {
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* part = partials + static_cast<size_t>(b) * slots * 16;
  double s = 0.0, q = 0.0;
  for (int k = lane; k < slots; k += 32) {
    s += static_cast<double>(part[k * 16 + 2 * warp]);
    q += static_cast<double>(part[k * 16 + 2 * warp + 1]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double mean = s / count;
    double var = q / count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    stats[b * GN_GROUPS + warp] = make_float2(static_cast<float>(mean),
                                              static_cast<float>(1.0 / sqrt(var + static_cast<double>(GN_EPS))));
  }
}

template <bool FAST> __device__ __forceinline__ float silu_f(float v) {
  if constexpr (FAST) {   // y*sigmoid(y) = h + h*tanh(h), h = y/2: one MUFU
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  }
  return v / (1.0f + expf(-v));
}

// y = (x - mean) * rstd * gamma + beta (+SiLU) -> padded T with circular halo.  4 x 8 channels per thread,
// all loads issued before the math.
template <typename T, bool IN_PADDED, bool SILU>
__global__ void __launch_bounds__(256) gn_apply_kernel(const void* __restrict__ in_, const float2* __restrict__ stats,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      int H, int W, int C, T* __restrict__ out) {
  constexpr bool FAST = sizeof(T) == 2;
  __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
  const int b = blockIdx.y;
  if (threadIdx.x < GN_GROUPS) {
    const float2 st = stats[b * GN_GROUPS + threadIdx.x];
    s_mean[threadIdx.x] = st.x;
    s_rstd[threadIdx.x] = st.y;
  }
  __syncthreads();
  const int cv = C / 8, cpg = C / GN_GROUPS;
  const int nvec = H * W * cv;
  const int Wp = W + 2, Hp = H + 2;
  const int e0 = blockIdx.x * 1024 + threadIdx.x;
  float v[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = e0 + k * 256;
    if (e >= nvec) continue;
    const int pix = e / cv, c = (e - pix * cv) * 8;
    if constexpr (IN_PADDED) {
      const int y = pix / W, x = pix - y * W;
      Vec8<T> iv;
      iv.load(static_cast<const T*>(in_) + ((static_cast<size_t>(b) * Hp + y + 1) * Wp + x + 1) * C + c);
      iv.get(v[k]);
    } else {
      Vec8<float> iv;
      iv.load(static_cast<const float*>(in_) + (static_cast<size_t>(b) * H * W + pix) * C + c);
      iv.get(v[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = e0 + k * 256;
    if (e >= nvec) continue;
    const int pix = e / cv, c = (e - pix * cv) * 8;
    const int y = pix / W, x = pix - y * W;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cpg;
      float yv = (v[k][j] - s_mean[g]) * s_rstd[g] * gm[j] + bt[j];
      if constexpr (SILU) yv = silu_f<FAST>(yv);
      v[k][j] = yv;
    }
    Vec8<T> ov;
    ov.set(v[k]);
    const int wy = halo_wrap(y, H), wx = halo_wrap(x, W);
    const size_t base = (static_cast<size_t>(b) * Hp + y + 1) * Wp + x + 1;
    ov.store(out + base * C + c);
    if (wy) ov.store(out + (base + static_cast<long long>(wy) * Wp) * C + c);
    if (wx) ov.store(out + (base + wx) * C + c);
    if (wy && wx) ov.store(out + (base + static_cast<long long>(wy) * Wp + wx) * C + c);
  }
}

int launch_gn_finalize(const float* partials, int slots, int B, int H, int W, int C, float2* stats, cudaStream_t st) {
  if (B <= 0) return TCS_OK;
  gn_finalize_kernel<<<B, 256, 0, st>>>(partials, slots, static_cast<double>(H) * W * (C / GN_GROUPS), stats);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

template <typename T>
int launch_gn_apply(const void* in, int in_padded, const float* partials, int slots, const float* gamma,
                    const float* beta, int B, int H, int W, int C, int silu, T* out, float2* stats, cudaStream_t st) {
  if (B <= 0) return TCS_OK;
  gn_finalize_kernel<<<B, 256, 0, st>>>(partials, slots, static_cast<double>(H) * W * (C / GN_GROUPS), stats);
  const int nvec = H * W * (C / 8);
  dim3 grid((nvec + 1023) / 1024, B);
  if (in_padded && !silu)
    gn_apply_kernel<T, true, false><<<grid, 256, 0, st>>>(in, stats, gamma, beta, H, W, C, out);
  else if (!in_padded && silu)
    gn_apply_kernel<T, false, true><<<grid, 256, 0, st>>>(in, stats, gamma, beta, H, W, C, out);
  else if (!in_padded && !silu)
    gn_apply_kernel<T, false, false><<<grid, 256, 0, st>>>(in, stats, gamma, beta, H, W, C, out);
  else
    gn_apply_kernel<T, true, true><<<grid, 256, 0, st>>>(in, stats, gamma, beta, H, W, C, out);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_gn_apply<float>(const void*, int, const float*, int, const float*, const float*, int, int, int, int, int, float*, float2*, cudaStream_t);
template int launch_gn_apply<__nv_bfloat16>(const void*, int, const float*, int, const float*, const float*, int, int, int, int, int, __nv_bfloat16*, float2*, cudaStream_t);

// GroupNorm of a whole 16x16x192 padded image in ONE kernel (attn.norm, sde_score_model.py:146-150): block = image,
// thread = (channel vector of 8, pixel lane); the image (96 KB in bf16) stays in registers between the statistics and
// the normalisation, so the three launches gn_stats -> gn_finalize -> gn_apply (and two re-reads) become one.
constexpr int GNI_LANES = 10;                      // pixel lanes: 24 channel vectors x 10 = 240 of 256 threads
constexpr int GNI_ITERS = (256 + GNI_LANES - 1) / GNI_LANES;
template <typename T>
__global__ void __launch_bounds__(256) gn_image16_kernel(const T* __restrict__ in, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, T* __restrict__ out) {
  constexpr int H = 16, W = 16, C = 192, CV = C / 8, Hp = H + 2, Wp = W + 2, CPG = C / GN_GROUPS;
  __shared__ float red[8][GN_GROUPS][2];
  __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
  const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int vc = t % CV, pl = t / CV;                // channel vector, pixel lane
  const bool active = pl < GNI_LANES;
  const int c = vc * 8, g = c / CPG;                 // 8 | 24: a vector never straddles a group
  const T* ib = in + static_cast<size_t>(b) * Hp * Wp * C + c;
  Vec8<T> v[GNI_ITERS];
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int k = 0; k < GNI_ITERS; ++k) {
    const int pix = pl + k * GNI_LANES;
    if (active && pix < H * W) {
      const int y = pix / W, x = pix - y * W;
      v[k].load(ib + (static_cast<size_t>(y + 1) * Wp + x + 1) * C);
    }
  }
#pragma unroll
  for (int k = 0; k < GNI_ITERS; ++k) {
    const int pix = pl + k * GNI_LANES;
    if (active && pix < H * W) {
      float f[8];
      v[k].get(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s += f[j]; q += f[j] * f[j]; }
    }
  }
  // block reduction per group: lanes of a warp may belong to different groups -> shared-memory atomics on 16 floats
  if (t < 8 * GN_GROUPS * 2) reinterpret_cast<float*>(red)[t] = 0.f;
  __syncthreads();
  if (active) {
    atomicAdd(&red[warp][g][0], s);
    atomicAdd(&red[warp][g][1], q);
  }
  __syncthreads();
  if (t < GN_GROUPS) {
    double sd = 0.0, qd = 0.0;
    for (int w = 0; w < 8; ++w) { sd += red[w][t][0]; qd += red[w][t][1]; }
    const double inv = 1.0 / (static_cast<double>(H) * W * CPG);
    const double mean = sd * inv;
    double var = qd * inv - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mean[t] = static_cast<float>(mean);
    s_rstd[t] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(GN_EPS)));
  }
  __syncthreads();
  (void)lane;
  if (!active) return;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
  const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const float mean = s_mean[g], rstd = s_rstd[g];
#pragma unroll
  for (int k = 0; k < GNI_ITERS; ++k) {
    const int pix = pl + k * GNI_LANES;
    if (pix >= H * W) continue;
    const int y = pix / W, x = pix - y * W;
    float f[8];
    v[k].get(f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * gm[j] + bt[j];
    Vec8<T> ov;
    ov.set(f);
    const int wy = halo_wrap(y, H), wx = halo_wrap(x, W);
    const size_t base = (static_cast<size_t>(b) * Hp + y + 1) * Wp + x + 1;
    ov.store(out + base * C + c);
    if (wy) ov.store(out + (base + static_cast<long long>(wy) * Wp) * C + c);
    if (wx) ov.store(out + (base + wx) * C + c);
    if (wy && wx) ov.store(out + (base + static_cast<long long>(wy) * Wp + wx) * C + c);
  }
}
template <typename T>
int launch_gn_image16(const T* in, int B, const float* gamma, const float* beta, T* out, cudaStream_t st) {
  if (B <= 0) return TCS_OK;
  gn_image16_kernel<T><<<B, 256, 0, st>>>(in, gamma, beta, out);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_gn_image16<float>(const float*, int, const float*, const float*, float*, cudaStream_t);
template int launch_gn_image16<__nv_bfloat16>(const __nv_bfloat16*, int, const float*, const float*, __nv_bfloat16*, cudaStream_t);

// statistics of a padded T tensor (one block per image; thread = channel) -> one partial slot
template <typename T>
__global__ void __launch_bounds__(192) gn_stats_kernel(const T* __restrict__ in, int H, int W, int C,
                                                      float* __restrict__ partials) {
  __shared__ float red[192][2];
  const int b = blockIdx.x, c = threadIdx.x;
  const int Wp = W + 2, Hp = H + 2;
  float s = 0.f, q = 0.f;
  if (c < C)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float v = to_f<T>(in[((static_cast<size_t>(b) * Hp + y + 1) * Wp + x + 1) * C + c]);
        s += v;
        q += v * v;
      }
  red[c][0] = s;
  red[c][1] = q;
  __syncthreads();
  if (c < GN_GROUPS) {
    const int cpg = C / GN_GROUPS;
    float ss = 0.f, qq = 0.f;
    for (int k = 0; k < cpg; ++k) { ss += red[c * cpg + k][0]; qq += red[c * cpg + k][1]; }
    partials[static_cast<size_t>(b) * 16 + 2 * c] = ss;
    partials[static_cast<size_t>(b) * 16 + 2 * c + 1] = qq;
  }
}
template <typename T>
int launch_gn_stats(const T* in, int B, int H, int W, int C, float* partials, cudaStream_t st) {
  if (B <= 0) return TCS_OK;
  if (C > 192) return fail(TCS_ERR_UNSUPPORTED, "gn_stats: C > 192");
  gn_stats_kernel<T><<<B, 192, 0, st>>>(in, H, W, C, partials);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_gn_stats<float>(const float*, int, int, int, int, float*, cudaStream_t);
template int launch_gn_stats<__nv_bfloat16>(const __nv_bfloat16*, int, int, int, int, float*, cudaStream_t);

// ------------------------------------------------------------------------------------------
// bilinear x2, align_corners=False, edge clamp (nn.Upsample, sde_score_model.py:217,221)
// value = wy0*(wx0*a + wx1*b) + wy1*(wx0*c + wx1*d)
// ------------------------------------------------------------------------------------------
// thread = (input pixel, 8 channels): reads the clamped 3x3 input neighbourhood once and produces the 2x2 output
// block (2iy..2iy+1, 2ix..2ix+1): horizontal lerps first, then vertical (the order PyTorch evaluates them in).
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_kernel(const T* __restrict__ in, int h, int w, int C,
                                                        T* __restrict__ out, long long total) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= total) return;
  const int cv = C / 8;
  const int c = static_cast<int>(e % cv) * 8;
  long long r = e / cv;
  const int ix = static_cast<int>(r % w); r /= w;
  const int iy = static_cast<int>(r % h);
  const int b = static_cast<int>(r / h);
  const int wp = w + 2, hp = h + 2;
  const int H = 2 * h, W = 2 * w, Wp = W + 2, Hp = H + 2;
  const int ys[3] = {max(iy - 1, 0), iy, min(iy + 1, h - 1)};
  const int xs[3] = {max(ix - 1, 0), ix, min(ix + 1, w - 1)};
  const T* ib = in + static_cast<size_t>(b) * hp * wp * C + c;
  float hl[3][8], hr[3][8];   // horizontal lerps for output columns 2ix (left) and 2ix+1 (right), per input row
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Vec8<T> v0, v1, v2;
    v0.load(ib + (static_cast<size_t>(ys[j] + 1) * wp + xs[0] + 1) * C);
    v1.load(ib + (static_cast<size_t>(ys[j] + 1) * wp + xs[1] + 1) * C);
    v2.load(ib + (static_cast<size_t>(ys[j] + 1) * wp + xs[2] + 1) * C);
    float f0[8], f1[8], f2[8];
    v0.get(f0); v1.get(f1); v2.get(f2);
    // output x = 2ix: src = ix - 0.25 -> (ix-1, ix) weights (0.25, 0.75), clamped at ix = 0 to (ix, ix) = f1
    // output x = 2ix+1: src = ix + 0.25 -> (ix, ix+1) weights (0.75, 0.25); the clamp is in xs[2]
    const float wl0 = ix == 0 ? 1.0f : 0.25f, wl1 = ix == 0 ? 0.0f : 0.75f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      hl[j][k] = ix == 0 ? f1[k] : (wl0 * f0[k] + wl1 * f1[k]);
      hr[j][k] = 0.75f * f1[k] + 0.25f * f2[k];
    }
  }
#pragma unroll
  for (int py = 0; py < 2; ++py) {
    const int y = 2 * iy + py;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const int x = 2 * ix + px;
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t0 = px ? hr[0][k] : hl[0][k], t1 = px ? hr[1][k] : hl[1][k], t2 = px ? hr[2][k] : hl[2][k];
        if (py == 0) o[k] = iy == 0 ? t1 : (0.25f * t0 + 0.75f * t1);
        else o[k] = 0.75f * t1 + 0.25f * t2;
      }
      Vec8<T> ov;
      ov.set(o);
      const int wy = halo_wrap(y, H), wx = halo_wrap(x, W);
      const size_t base = (static_cast<size_t>(b) * Hp + y + 1) * Wp + x + 1;
      ov.store(out + base * C + c);
      if (wy) ov.store(out + (base + static_cast<long long>(wy) * Wp) * C + c);
      if (wx) ov.store(out + (base + wx) * C + c);
      if (wy && wx) ov.store(out + (base + static_cast<long long>(wy) * Wp + wx) * C + c);
    }
  }
}
// Tiled variant: a block stages UP_BR + 2 (edge-clamped) input rows of one image in shared memory, so every input value
// is read from L2 1.5 times instead of 9 (the plain kernel was bound by its L2 read traffic: 9 x 16 B per thread).
// A thread owns one (input column, 8 channels) item and walks down the staged rows: the horizontal lerps of a row are
// computed once and kept for the next row (the first tiled version recomputed them for each of the three output rows that
// use them and was ISSUE-bound: 80 % issue-active, ALU pipe 66 %, 3.7 TB/s), the arithmetic is packed (FMUL2 / FFMA2).
constexpr int UP_BR = 4;
constexpr int UP_THREADS = 192;          // = w * (C / 2) / 8 for both upsamples of the network: one item per thread, a block owns one channel half
__device__ __forceinline__ void pk_mul2(float& d0, float& d1, float a0, float a1, float b) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b));
}
__device__ __forceinline__ void pk_fma2(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
// o = wa * a + wb * b over 8 channels
__device__ __forceinline__ void lerp8(float* o, float wa, const float* a, float wb, const float* b) {
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    float t0, t1;
    pk_mul2(t0, t1, a[k], a[k + 1], wa);
    pk_fma2(o[k], o[k + 1], b[k], b[k + 1], wb, t0, t1);
  }
}
// store 8 channels of output pixel (y, x) at dst plus its circular-halo copies (rare: border pixels only).
// ey / ex = -1, 0, +1: the pixel is on the last / no / the first row (column); the copy sits H rows (W columns) away.
template <typename T>
__device__ __forceinline__ void up_emit(T* __restrict__ dst, const float* o, int ey, int ex, int H, int W, int C) {
  Vec8<T> ov;
  ov.set(o);
  ov.store(dst);
  if (ey | ex) {
    const long long wy = static_cast<long long>(ey) * H * (W + 2) * C, wx = static_cast<long long>(ex) * W * C;
    if (ey) ov.store(dst + wy);
    if (ex) ov.store(dst + wx);
    if (ey && ex) ov.store(dst + wy + wx);
  }
}
template <typename T>
__global__ void __launch_bounds__(UP_THREADS, 4) upsample2x_tiled_kernel(const T* __restrict__ in, int h, int w, int C,
                                                                         T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char up_smem[];
  T* tile = reinterpret_cast<T*>(up_smem);            // [UP_BR + 2][w][C / 2]: this block's channel half
  const int bands = h / UP_BR;
  const int chalf = blockIdx.x & 1, bb = blockIdx.x >> 1;
  const int b = bb / bands, iy0 = (bb - b * bands) * UP_BR;
  const int Ch = C / 2, cv = Ch / 8, wp = w + 2, hp = h + 2;
  const T* ib = in + static_cast<size_t>(b) * hp * wp * C + chalf * Ch;
  out += chalf * Ch;
  const int row_vecs = w * cv;
  for (int e = threadIdx.x; e < (UP_BR + 2) * row_vecs; e += UP_THREADS) {
    const int j = e / row_vecs, r = e - j * row_vecs;
    const int x = r / cv, k = r - x * cv;                       // pixel, channel vector inside the half
    const int ys = min(max(iy0 - 1 + j, 0), h - 1);
    Vec8<T> v;
    v.load(ib + (static_cast<size_t>(ys + 1) * wp + 1 + x) * C + k * 8);
    v.store(tile + static_cast<size_t>(e) * 8);
  }
  __syncthreads();
  const int H = 2 * h, W = 2 * w;
  const long long rs = static_cast<long long>(W + 2) * C;          // output row stride in elements
  for (int r = threadIdx.x; r < row_vecs; r += UP_THREADS) {
    const int ix = r / cv, c = (r - ix * cv) * 8;
    const int x0 = max(ix - 1, 0), x2 = min(ix + 1, w - 1);
    // output pixel (y = 2 iy0, x = 2 ix) in the padded tensor; the column halo copies sit W pixels to the right / left
    T* const o00 = out + ((static_cast<size_t>(b) * (H + 2) + 2 * iy0 + 1) * (W + 2) + 2 * ix + 1) * C + c;
    const int exl = ix == 0 ? 1 : 0, exr = ix == w - 1 ? -1 : 0;       // x = 0 -> also column W; x = W - 1 -> also column -1
    float pl[8], pr[8];      // horizontal lerps of the previous staged row: output columns 2ix (left) and 2ix+1 (right)
#pragma unroll
    for (int j = 0; j < UP_BR + 2; ++j) {            // staged row j = input row iy0 - 1 + j (edge-clamped)
      const T* trow = tile + (static_cast<size_t>(j) * w) * Ch + c;
      Vec8<T> v0, v1, v2;
      v0.load(trow + static_cast<size_t>(x0) * Ch);
      v1.load(trow + static_cast<size_t>(ix) * Ch);
      v2.load(trow + static_cast<size_t>(x2) * Ch);
      float f0[8], f1[8], f2[8], cl[8], cr[8];
      v0.get(f0); v1.get(f1); v2.get(f2);
      lerp8(cl, 0.25f, f0, 0.75f, f1);
      if (ix == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) cl[k] = f1[k];
      }
      lerp8(cr, 0.25f, f2, 0.75f, f1);
      if (j >= 2) {                                   // odd output row of input row iy0 + j - 2: 0.75 prev + 0.25 cur
        T* const dst = o00 + (2 * (j - 2) + 1) * rs;
        const int ey = (iy0 + j - 2 == h - 1) ? -1 : 0;                     // y = H - 1 -> also row -1
        float o[8];
        lerp8(o, 0.25f, cl, 0.75f, pl);
        up_emit<T>(dst, o, ey, exl, H, W, C);
        lerp8(o, 0.25f, cr, 0.75f, pr);
        up_emit<T>(dst + C, o, ey, exr, H, W, C);
      }
      if (j >= 1 && j <= UP_BR) {                     // even output row of input row iy0 + j - 1: 0.25 prev + 0.75 cur
        const bool top = iy0 + j - 1 == 0;
        T* const dst = o00 + (2 * (j - 1)) * rs;
        const int ey = top ? 1 : 0;                                         // y = 0 -> also row H
        float o[8];
        lerp8(o, 0.25f, pl, 0.75f, cl);
        if (top) {
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = cl[k];
        }
        up_emit<T>(dst, o, ey, exl, H, W, C);
        lerp8(o, 0.25f, pr, 0.75f, cr);
        if (top) {
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = cr[k];
        }
        up_emit<T>(dst + C, o, ey, exr, H, W, C);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { pl[k] = cl[k]; pr[k] = cr[k]; }
    }
  }
}

template <typename T>
int launch_upsample2x(const T* in, int B, int h, int w, int C, T* out, cudaStream_t st) {
  if (B <= 0) return TCS_OK;
  const size_t tile_bytes = static_cast<size_t>(UP_BR + 2) * w * (C / 2) * sizeof(T);
  if (h % UP_BR == 0 && C % 16 == 0 && tile_bytes <= 48 * 1024) {
    upsample2x_tiled_kernel<T><<<static_cast<unsigned>(B * (h / UP_BR) * 2), UP_THREADS, tile_bytes, st>>>(in, h, w, C, out);
    TCS_CUDA(cudaGetLastError());
    return TCS_OK;
  }
  const long long total = static_cast<long long>(B) * h * w * (C / 8);
  upsample2x_kernel<T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, h, w, C, out, total);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_upsample2x<float>(const float*, int, int, int, int, float*, cudaStream_t);
template int launch_upsample2x<__nv_bfloat16>(const __nv_bfloat16*, int, int, int, int, __nv_bfloat16*, cudaStream_t);

// ------------------------------------------------------------------------------------------
// attention: 4 heads x 256 tokens x d=48 (SelfAttention2d, sde_score_model.py:146-164)
// block = (image, head); thread = query token; K,V in shared memory; online softmax
// ------------------------------------------------------------------------------------------
constexpr int ATT_TOK = 256, ATT_D = 48, ATT_C = 192;
template <typename T>
__global__ void __launch_bounds__(ATT_TOK) attention_kernel(const T* __restrict__ qkv, T* __restrict__ yout) {
  extern __shared__ float att_smem[];
  float* Ks = att_smem;
  float* Vs = att_smem + ATT_TOK * ATT_D;
  const int b = blockIdx.x >> 2, head = blockIdx.x & 3, t = threadIdx.x;
  const T* base = qkv + static_cast<size_t>(b) * ATT_TOK * (3 * ATT_C);
  for (int e = t; e < ATT_TOK * (ATT_D / 8); e += ATT_TOK) {
    const int tok = e / (ATT_D / 8), d8 = (e % (ATT_D / 8)) * 8;
    Vec8<T> kv, vv;
    kv.load(base + static_cast<size_t>(tok) * (3 * ATT_C) + ATT_C + head * ATT_D + d8);
    vv.load(base + static_cast<size_t>(tok) * (3 * ATT_C) + 2 * ATT_C + head * ATT_D + d8);
    kv.get(Ks + tok * ATT_D + d8);
    vv.get(Vs + tok * ATT_D + d8);
  }
  float q[ATT_D], acc[ATT_D];
  const float scale = 0.14433756729740643f;  // 1/sqrt(48)
#pragma unroll
  for (int d8 = 0; d8 < ATT_D; d8 += 8) {
    Vec8<T> qv;
    qv.load(base + static_cast<size_t>(t) * (3 * ATT_C) + head * ATT_D + d8);
    qv.get(q + d8);
  }
#pragma unroll
  for (int d = 0; d < ATT_D; ++d) { q[d] *= scale; acc[d] = 0.f; }
  __syncthreads();
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < ATT_TOK; j0 += 8) {
    float s[8];
    float mx = m;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float4* kr = reinterpret_cast<const float4*>(Ks + (j0 + jj) * ATT_D);
      float a = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < ATT_D / 4; ++d4) {
        const float4 k4 = kr[d4];
        a = fmaf(q[4 * d4], k4.x, a); a = fmaf(q[4 * d4 + 1], k4.y, a);
        a = fmaf(q[4 * d4 + 2], k4.z, a); a = fmaf(q[4 * d4 + 3], k4.w, a);
      }
      s[jj] = a;
      mx = fmaxf(mx, a);
    }
    const float corr = expf(m - mx);
    m = mx;
    l *= corr;
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) acc[d] *= corr;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float pj = expf(s[jj] - m);
      l += pj;
      const float4* vr = reinterpret_cast<const float4*>(Vs + (j0 + jj) * ATT_D);
#pragma unroll
      for (int d4 = 0; d4 < ATT_D / 4; ++d4) {
        const float4 v4 = vr[d4];
        acc[4 * d4] = fmaf(pj, v4.x, acc[4 * d4]); acc[4 * d4 + 1] = fmaf(pj, v4.y, acc[4 * d4 + 1]);
        acc[4 * d4 + 2] = fmaf(pj, v4.z, acc[4 * d4 + 2]); acc[4 * d4 + 3] = fmaf(pj, v4.w, acc[4 * d4 + 3]);
      }
    }
  }
  const float inv = 1.0f / l;
  T* orow = yout + (static_cast<size_t>(b) * ATT_TOK + t) * ATT_C + head * ATT_D;
#pragma unroll
  for (int d8 = 0; d8 < ATT_D; d8 += 8) {
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = acc[d8 + k] * inv;
    Vec8<T> ov;
    ov.set(o);
    ov.store(orow + d8);
  }
}
// (The bf16 mma.sync version of this kernel is gone: the tcgen05 engine runs the whole attention block in attn_tc.cu;
// this CUDA-core kernel serves the fp32 mode, the CUDA-core cross-check engine and the attn.qkv / attn.y debug taps.)
template <typename T>
int launch_attention(const T* qkv, int B, T* y, cudaStream_t st) {
  if (B <= 0) return TCS_OK;
  const size_t smem = 2 * ATT_TOK * ATT_D * sizeof(float);
  static bool done = false;
  if (!done) {
    TCS_CUDA(cudaFuncSetAttribute(attention_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    done = true;
  }
  attention_kernel<T><<<B * N_HEADS, ATT_TOK, smem, st>>>(qkv, y);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_attention<float>(const float*, int, float*, cudaStream_t);
template int launch_attention<__nv_bfloat16>(const __nv_bfloat16*, int, __nv_bfloat16*, cudaStream_t);

// ------------------------------------------------------------------------------------------
// generic SIMT convolution: tile = 64 pixels x 96 output channels, 256 threads (4 px x 6 oc each)
// ------------------------------------------------------------------------------------------
struct SimtConvParams {
  int B, H, W, HW;            // output
  int ksize, stride, nsrc;
  int csrc[2], Hin[2], Win[2], base_off[2];
  int cin_tot, ntot;
  EpiArgs epi;
};

template <typename T, int EPI>
__global__ void __launch_bounds__(256) conv_simt_kernel(const T* __restrict__ src0, const T* __restrict__ src1,
                                                       const float* __restrict__ wp, const SimtConvParams p) {
  __shared__ __align__(16) float As[32][68];
  __shared__ __align__(16) float Bs[32][96];
  __shared__ float red[256][2];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.x * 64, n_off = blockIdx.y * 96;
  const int b = m0 / p.HW;
  // A-load role: pixel lp = t/4, 8 channels at (t%4)*8
  const int lp = t >> 2, lc = (t & 3) * 8;
  const int lrem = m0 + lp - b * p.HW;
  const int ly = lrem / p.W, lx = lrem - ly * p.W;
  float acc[4][6];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < p.ksize * p.ksize; ++tap) {
    const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
    int coff = 0;
    for (int s = 0; s < p.nsrc; ++s) {
      const T* src = s == 0 ? src0 : src1;
      const int C = p.csrc[s];
      const int iy = ly * p.stride + ky + p.base_off[s], ix = lx * p.stride + kx + p.base_off[s];
      const T* arow = src + ((static_cast<size_t>(b) * p.Hin[s] + iy) * p.Win[s] + ix) * C + lc;
      for (int cb = 0; cb < C; cb += 32) {
        Vec8<T> av;
        av.load(arow + cb);
        float f[8];
        av.get(f);
#pragma unroll
        for (int k = 0; k < 8; ++k) As[lc + k][lp] = f[k];
        const float* wrow = wp + (static_cast<size_t>(tap) * p.cin_tot + coff + cb) * p.ntot + n_off;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int e = t + r * 256;           // 768 float4 = 32 rows x 24
          const int kk = e / 24, c4 = (e - kk * 24) * 4;
          *reinterpret_cast<float4*>(&Bs[kk][c4]) = __ldg(reinterpret_cast<const float4*>(wrow + static_cast<size_t>(kk) * p.ntot + c4));
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
          const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
          const float2 b0 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 6]);
          const float2 b1 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 6 + 2]);
          const float2 b2 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 6 + 4]);
          const float av4[4] = {a4.x, a4.y, a4.z, a4.w};
          const float bv[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(av4[i], bv[j], acc[i][j]);
        }
        __syncthreads();
      }
      coff += C;
    }
  }

  // ---- epilogue -------------------------------------------------------------------------
  const int oc0 = n_off + tx * 6;
  float bias[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) bias[j] = __ldg(p.epi.bias + oc0 + j);
  if constexpr (EPI == EPI_RAW_STATS) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      float* o = static_cast<float*>(p.epi.out) + static_cast<size_t>(m) * p.epi.ldo + oc0;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float v = acc[i][j] + bias[j];
        o[j] = v;
        s += v;
        q += v * v;
      }
    }
    red[t][0] = s;
    red[t][1] = q;
    __syncthreads();
    const int cpg = p.ntot / GN_GROUPS;      // 12 or 24
    const int gpt = 96 / cpg;                // groups in this n-tile: 8 or 4
    if (t < gpt) {
      const int txpg = cpg / 6;              // tx values per group: 2 or 4
      float ss = 0.f, qq = 0.f;
      for (int yy = 0; yy < 16; ++yy)
        for (int xx = 0; xx < txpg; ++xx) { ss += red[yy * 16 + t * txpg + xx][0]; qq += red[yy * 16 + t * txpg + xx][1]; }
      const int slot = (m0 - b * p.HW) / 64;
      const int g = n_off / cpg + t;
      float* dst = p.epi.partials + (static_cast<size_t>(b) * p.epi.slots + slot) * 16 + 2 * g;
      dst[0] = ss;
      dst[1] = qq;
    }
  } else if constexpr (EPI == EPI_PADDED) {
    const int Wp = p.W + 2, Hp = p.H + 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rem = m0 + ty * 4 + i - b * p.HW;
      const int y = rem / p.W, x = rem - y * p.W;
      const size_t pix = (static_cast<size_t>(b) * Hp + y + 1) * Wp + x + 1;
      T ov[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        float v = acc[i][j] + bias[j];
        if (p.epi.residual) v += to_f<T>(static_cast<const T*>(p.epi.residual)[pix * p.ntot + oc0 + j]);
        ov[j] = from_f<T>(v);
      }
      const int wy = halo_wrap(y, p.H), wx = halo_wrap(x, p.W);
      T* ob = static_cast<T*>(p.epi.out);
#pragma unroll
      for (int cy = 0; cy < 2; ++cy) {
        if (cy && !wy) continue;
#pragma unroll
        for (int cx = 0; cx < 2; ++cx) {
          if (cx && !wx) continue;
          T* o = ob + (pix + static_cast<long long>(cy ? wy : 0) * Wp + (cx ? wx : 0)) * p.epi.ldo + oc0;
#pragma unroll
          for (int j = 0; j < 6; ++j) o[j] = ov[j];
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      T* o = static_cast<T*>(p.epi.out) + static_cast<size_t>(m) * p.epi.ldo + oc0;
#pragma unroll
      for (int j = 0; j < 6; ++j) o[j] = from_f<T>(acc[i][j] + bias[j]);
    }
  }
}

void conv_simt_pack_weights(const ConvGeom& g, const float* w, float* out, bool round_bf16) {
  const int k = g.ksize;
  int cin = 0;
  for (int s = 0; s < g.nsrc; ++s) cin += g.csrc[s];
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx)
      for (int c = 0; c < cin; ++c)
        for (int n = 0; n < g.ntot; ++n) {
          float v = w[((static_cast<size_t>(n) * cin + c) * k + ky) * k + kx];
          if (round_bf16) v = __bfloat162float(__float2bfloat16(v));
          out[((static_cast<size_t>(ky) * k + kx) * cin + c) * g.ntot + n] = v;
        }
}

template <typename T>
int launch_conv_simt(const ConvGeom& g, const T* src0, const T* src1, const float* wpacked, int epi, const EpiArgs& ea,
                     cudaStream_t st) {
  if (g.B <= 0) return TCS_OK;
  SimtConvParams p;
  p.B = g.B; p.H = g.H; p.W = g.W; p.HW = g.H * g.W;
  p.ksize = g.ksize; p.stride = g.stride; p.nsrc = g.nsrc;
  p.cin_tot = 0;
  const int conv_pad = g.ksize == 1 ? 0 : 1;
  for (int s = 0; s < 2; ++s) {
    const int si = s < g.nsrc ? s : 0;
    p.csrc[s] = g.csrc[si];
    p.Hin[s] = g.H * g.stride + 2 * g.in_pad[si];
    p.Win[s] = g.W * g.stride + 2 * g.in_pad[si];
    p.base_off[s] = g.in_pad[si] - conv_pad;
    if (p.base_off[s] < 0) return fail(TCS_ERR_BAD_ARGUMENT, "conv_simt: a 3x3/4x4 conv needs a padded source");
    if (s < g.nsrc) p.cin_tot += g.csrc[s];
    if (g.csrc[si] % 32) return fail(TCS_ERR_UNSUPPORTED, "conv_simt: C_in must be a multiple of 32");
  }
  p.ntot = g.ntot;
  p.epi = ea;
  if (p.HW % 64 || g.ntot % 96) return fail(TCS_ERR_UNSUPPORTED, "conv_simt: needs HW%64==0 and C_out%96==0");
  dim3 grid(g.B * p.HW / 64, g.ntot / 96);
  switch (epi) {
    case EPI_RAW_STATS: conv_simt_kernel<T, EPI_RAW_STATS><<<grid, 256, 0, st>>>(src0, src1, wpacked, p); break;
    case EPI_PADDED: conv_simt_kernel<T, EPI_PADDED><<<grid, 256, 0, st>>>(src0, src1, wpacked, p); break;
    case EPI_PLAIN: conv_simt_kernel<T, EPI_PLAIN><<<grid, 256, 0, st>>>(src0, src1, wpacked, p); break;
    default: return fail(TCS_ERR_BAD_ARGUMENT, "conv_simt: bad epilogue");
  }
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_conv_simt<float>(const ConvGeom&, const float*, const float*, const float*, int, const EpiArgs&, cudaStream_t);
template int launch_conv_simt<__nv_bfloat16>(const ConvGeom&, const __nv_bfloat16*, const __nv_bfloat16*, const float*, int, const EpiArgs&, cudaStream_t);

// ------------------------------------------------------------------------------------------
// out conv 96 -> 1 (3x3 circular) + CFG combine  (sde_score_model.py:225,266 and :418-423)
// 8 lanes per output pixel, 12 channels per lane; 32 pixels per 256-thread block
// ------------------------------------------------------------------------------------------
// lane `sl` of the 8 lanes that share a pixel handles the 12 channels 12 sl .. 12 sl + 11 (three 4-channel loads)
template <typename T> __device__ __forceinline__ void load4(const T* p, float* f);
template <> __device__ __forceinline__ void load4<float>(const float* p, float* f) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  f[0] = a.x; f[1] = a.y; f[2] = c.x; f[3] = c.y;
}

// Each thread walks down OUT_ROWS + 2 input rows with three running accumulators (the outputs that see the current input
// row as tap ky = 2, 1, 0): an input row is loaded once per block instead of three times.  The kernel is bound by
// L2 -> SM traffic (9 taps x 96 channels per output pixel): 3.2 ms -> ~1 ms per 2048 fp32 images.  Every output still
// accumulates in the order (ky, kx, channel), so results do not depend on OUT_ROWS.
constexpr int OUT_ROWS = 8;
template <typename T>
__global__ void __launch_bounds__(256, 2) out_conv_kernel(const T* __restrict__ act, const float* __restrict__ w, float bias,
                                                         int dup, float guidance, float* __restrict__ eps) {
  __shared__ __align__(16) float ws[9 * 96];
  for (int e = threadIdx.x; e < 9 * 96; e += 256) ws[e] = w[e];
  __syncthreads();
  const int sl = threadIdx.x & 7;                       // channel slice: 12 channels
  constexpr int BLK_PER_IMG = (IMG / OUT_ROWS) * (IMG / 32);
  const int i = blockIdx.x / BLK_PER_IMG, rb = blockIdx.x % BLK_PER_IMG;
  const int y0 = (rb >> 1) * OUT_ROWS, x = (rb & 1) * 32 + (threadIdx.x >> 3);
  constexpr int Wp = IMG + 2;
  const size_t img_stride = static_cast<size_t>(Wp) * Wp * 96;
  const T* base = act + static_cast<size_t>(i) * dup * img_stride + sl * 12;
  float acc[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};   // [branch][ky = 2, 1, 0]
#pragma unroll 1
  for (int ry = 0; ry < OUT_ROWS + 2; ++ry) {            // padded input row y0 + ry
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u >= dup) break;
      const T* row = base + u * img_stride + (static_cast<size_t>(y0 + ry) * Wp + x) * 96;
#pragma unroll 1
      for (int kx = 0; kx < 3; ++kx) {
        float f[12];
#pragma unroll
        for (int j = 0; j < 3; ++j) load4<T>(row + kx * 96 + j * 4, f + j * 4);
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) {                 // accumulator s3 sees this row as tap ky = 2 - s3
          const float* wk = ws + ((2 - s3) * 3 + kx) * 96 + sl * 12;
#pragma unroll
          for (int j = 0; j < 12; j += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wk + j);
            acc[u][s3] = fmaf(f[j], w4.x, acc[u][s3]); acc[u][s3] = fmaf(f[j + 1], w4.y, acc[u][s3]);
            acc[u][s3] = fmaf(f[j + 2], w4.z, acc[u][s3]); acc[u][s3] = fmaf(f[j + 3], w4.w, acc[u][s3]);
          }
        }
      }
    }
    // output row ry - 2 has now seen its three input rows
    float e[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float v = acc[u][0];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      e[u] = v + bias;
      acc[u][0] = acc[u][1]; acc[u][1] = acc[u][2]; acc[u][2] = 0.f;
    }
    if (ry >= 2 && sl == 0)
      eps[static_cast<size_t>(i) * IMG_PIX + (y0 + ry - 2) * IMG + x] = dup == 2 ? e[1] + guidance * (e[0] - e[1]) : e[0];
  }
}
template <typename T>
int launch_out_conv(const T* act, const float* w, float bias, int n, int dup, float guidance, float* eps,
                    cudaStream_t st) {
  if (n <= 0) return TCS_OK;
  out_conv_kernel<T><<<n * (IMG / OUT_ROWS) * (IMG / 32), 256, 0, st>>>(act, w, bias, dup, guidance, eps);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_out_conv<float>(const float*, const float*, float, int, int, float, float*, cudaStream_t);
template int launch_out_conv<__nv_bfloat16>(const __nv_bfloat16*, const float*, float, int, int, float, float*, cudaStream_t);

// ------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pad_from_plain_kernel(const float* __restrict__ in, int H, int W, int C, int pad,
                                                            T* __restrict__ out, long long total) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= total) return;
  const int c = static_cast<int>(e % C);
  long long r = e / C;
  const int Wp = W + 2 * pad, Hp = H + 2 * pad;
  const int px = static_cast<int>(r % Wp); r /= Wp;
  const int py = static_cast<int>(r % Hp);
  const long long b = r / Hp;
  const int y = (py - pad + H) % H, x = (px - pad + W) % W;
  out[e] = from_f<T>(in[((b * H + y) * W + x) * C + c]);
}
template <typename T>
int launch_pad_from_plain(const float* in, int B, int H, int W, int C, int pad, T* out, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * (H + 2 * pad) * (W + 2 * pad) * C;
  if (total <= 0) return TCS_OK;
  pad_from_plain_kernel<T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, H, W, C, pad, out, total);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_pad_from_plain<float>(const float*, int, int, int, int, int, float*, cudaStream_t);
template int launch_pad_from_plain<__nv_bfloat16>(const float*, int, int, int, int, int, __nv_bfloat16*, cudaStream_t);

template <typename T>
__global__ void __launch_bounds__(256) unpad_kernel(const T* __restrict__ in, int H, int W, int C, int pad,
                                                   float* __restrict__ out, long long total) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= total) return;
  const int c = static_cast<int>(e % C);
  long long r = e / C;
  const int x = static_cast<int>(r % W); r /= W;
  const int y = static_cast<int>(r % H);
  const long long b = r / H;
  out[e] = to_f<T>(in[((b * (H + 2 * pad) + y + pad) * (W + 2 * pad) + x + pad) * C + c]);
}
template <typename T>
int launch_unpad_to_f32(const T* in, int B, int H, int W, int C, int pad, float* out, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * H * W * C;
  if (total <= 0) return TCS_OK;
  unpad_kernel<T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, H, W, C, pad, out, total);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}
template int launch_unpad_to_f32<float>(const float*, int, int, int, int, int, float*, cudaStream_t);
template int launch_unpad_to_f32<__nv_bfloat16>(const __nv_bfloat16*, int, int, int, int, int, float*, cudaStream_t);

}  // namespace tcs
