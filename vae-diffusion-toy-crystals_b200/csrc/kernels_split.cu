// kernels_split.cu — glue kernels of the fp32 mode on the tensor pipe ("bf16x3", precision = fp32 with engine = tcgen05).
//
// north_star's fp32 mode asks for per-evaluation eps within 1e-4 of the reference (sde_score_model.py:243-266 in fp32).
// kind::tf32 (10-bit mantissa) cannot hold that through ~20 layers, so the convolutions run on the bf16 tensor pipe with
// every fp32 operand written as a pair  a = a_hi + a_lo,  a_hi = bf16(a),  a_lo = bf16(a - a_hi)  and
//     a * w  ~=  a_hi*w_hi + a_hi*w_lo + a_lo*w_hi        (fp32 accumulation in TMEM)
// which conv_tc_kernel executes as a K-tripled implicit GEMM over the sources [a_hi, a_hi, a_lo] against the packed
// weight segments [w_hi, w_lo, w_hi] (conv_tc.cu).  The dropped a_lo*w_lo term and the residual of the two-term
// representation are ~2^-17 relative per product.  Everything that is not a convolution stays in fp32 (GroupNorm
// statistics, SiLU with expf, attention, upsample, first and last conv: the FFMA kernels of kernels_simt.cu).
//
// A "split tensor" is ONE allocation of 4 bytes per element: the bf16 hi plane first, the bf16 lo plane `lo_off`
// elements later (the planes have the layout of the fp32 tensor they replace: padded NHWC with circular halo, or plain).
#include "kernels.cuh"

namespace tcs {

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(v);
  lo = __float2bfloat16(v - __bfloat162float(hi));
}
struct Split8 {   // 8 consecutive channels as hi / lo bf16 (16 bytes each)
  uint4 hi, lo;
  __device__ __forceinline__ void set(const float* f) {
    __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_bf16(f[i], h[i], l[i]);
    hi = *reinterpret_cast<const uint4*>(h);
    lo = *reinterpret_cast<const uint4*>(l);
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p, size_t lo_off) const {
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + lo_off) = lo;
  }
};
__device__ __forceinline__ int wrap_off(int v, int n) { return v == 0 ? n : (v == n - 1 ? -n : 0); }
__device__ __forceinline__ void load8(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---- fp32 tensor (any layout, n8 groups of 8 elements) -> split tensor of the same layout ------------------------------
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                   size_t lo_off, size_t n8) {
  const size_t i = blockIdx.x * 256ull + threadIdx.x;
  if (i >= n8) return;
  float f[8];
  load8(in + i * 8, f);
  Split8 s;
  s.set(f);
  s.store(out + i * 8, lo_off);
}
int launch_split(const float* in, size_t elems, __nv_bfloat16* out, size_t lo_off, cudaStream_t st) {
  if (elems == 0) return TCS_OK;
  if (elems % 8) return fail(TCS_ERR_BAD_ARGUMENT, "launch_split: element count must be a multiple of 8");
  const size_t n8 = elems / 8;
  split_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, st>>>(in, out, lo_off, n8);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

// ---- raw fp32 conv output [B,H,W,C] (bias included) -> padded tensor with circular halo ---------------------------------
// MODE 0: GroupNorm(stats) * gamma + beta, SiLU (exact expf form)      -> split tensor      (conv -> GN -> SiLU -> next conv)
// MODE 1: identity                                                      -> split tensor      (ds / us convs -> next conv)
// MODE 2: + residual (padded fp32)                                      -> padded fp32       (attn.proj + x -> upsample)
template <int MODE>
__global__ void __launch_bounds__(256) raw_to_padded_kernel(const float* __restrict__ raw, const float2* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ residual, int H, int W, int C,
                                                           void* __restrict__ out_, size_t lo_off) {
  __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
  const int b = blockIdx.y;
  if (MODE == 0) {
    if (threadIdx.x < GN_GROUPS) {
      const float2 st = stats[b * GN_GROUPS + threadIdx.x];
      s_mean[threadIdx.x] = st.x;
      s_rstd[threadIdx.x] = st.y;
    }
    __syncthreads();
  }
  const int cv = C / 8, cpg = C / GN_GROUPS;
  const int nvec = H * W * cv;
  const int Wp = W + 2, Hp = H + 2;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= nvec) return;
  const int pix = e / cv, c = (e - pix * cv) * 8;
  const int y = pix / W, x = pix - y * W;
  float v[8];
  load8(raw + (static_cast<size_t>(b) * H * W + pix) * C + c, v);
  const size_t base = (static_cast<size_t>(b) * Hp + y + 1) * Wp + x + 1;
  if (MODE == 0) {
    float gm[8], bt[8];
    load8(gamma + c, gm);
    load8(beta + c, bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cpg;
      const float yv = (v[j] - s_mean[g]) * s_rstd[g] * gm[j] + bt[j];
      v[j] = yv / (1.0f + expf(-yv));
    }
  } else if (MODE == 2) {
    float r[8];
    load8(residual + base * C + c, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += r[j];
  }
  const int wy = wrap_off(y, H), wx = wrap_off(x, W);
  if (MODE == 2) {
    float* out = static_cast<float*>(out_);
    auto st8 = [&](size_t px) {
      *reinterpret_cast<float4*>(out + px * C + c) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(out + px * C + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
    };
    st8(base);
    if (wy) st8(base + static_cast<long long>(wy) * Wp);
    if (wx) st8(base + wx);
    if (wy && wx) st8(base + static_cast<long long>(wy) * Wp + wx);
  } else {
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(out_);
    Split8 s;
    s.set(v);
    s.store(out + base * C + c, lo_off);
    if (wy) s.store(out + (base + static_cast<long long>(wy) * Wp) * C + c, lo_off);
    if (wx) s.store(out + (base + wx) * C + c, lo_off);
    if (wy && wx) s.store(out + (base + static_cast<long long>(wy) * Wp + wx) * C + c, lo_off);
  }
}

int launch_raw_to_padded(int mode, const float* raw, const float2* stats, const float* gamma, const float* beta,
                         const float* residual, int B, int H, int W, int C, void* out, size_t lo_off, cudaStream_t st) {
  if (B <= 0) return TCS_OK;
  const int nvec = H * W * (C / 8);
  dim3 grid((nvec + 255) / 256, B);
  if (mode == 0) raw_to_padded_kernel<0><<<grid, 256, 0, st>>>(raw, stats, gamma, beta, residual, H, W, C, out, lo_off);
  else if (mode == 1) raw_to_padded_kernel<1><<<grid, 256, 0, st>>>(raw, stats, gamma, beta, residual, H, W, C, out, lo_off);
  else raw_to_padded_kernel<2><<<grid, 256, 0, st>>>(raw, stats, gamma, beta, residual, H, W, C, out, lo_off);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

// ---- split padded tensor -> fp32 NHWC [B,H,W,C] (debug taps) --------------------------------------------------------------
__global__ void __launch_bounds__(256) unsplit_kernel(const __nv_bfloat16* __restrict__ in, size_t lo_off, int H, int W, int C,
                                                     int pad, float* __restrict__ out, long long total) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= total) return;
  const int c = static_cast<int>(e % C);
  long long r = e / C;
  const int x = static_cast<int>(r % W); r /= W;
  const int y = static_cast<int>(r % H);
  const long long b = r / H;
  const size_t idx = ((b * (H + 2 * pad) + y + pad) * (W + 2 * pad) + x + pad) * C + c;
  out[e] = __bfloat162float(in[idx]) + __bfloat162float(in[idx + lo_off]);
}
int launch_unsplit_to_f32(const __nv_bfloat16* in, size_t lo_off, int B, int H, int W, int C, int pad, float* out,
                          cudaStream_t st) {
  const long long total = static_cast<long long>(B) * H * W * C;
  if (total <= 0) return TCS_OK;
  unsplit_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, lo_off, H, W, C, pad, out, total);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

}  // namespace tcs
