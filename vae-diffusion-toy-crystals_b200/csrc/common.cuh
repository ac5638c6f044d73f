// common.cuh — shared declarations for libtcs (B200 / sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/tcs.h"

namespace tcs {

constexpr int IMG = 64;            // the sampler hard-codes 64x64 (sde_score_model.py:329,340)
constexpr int IMG_PIX = IMG * IMG;
constexpr int GN_GROUPS = 8;       // _gn_groups(96) == _gn_groups(192) == 8 (sde_score_model.py:89-94)
constexpr float GN_EPS = 1e-5f;
constexpr int N_HEADS = 4;

// thread-local error plumbing (tcs_api.cu)
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define TCS_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::tcs::fail(TCS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
  } while (0)

#define TCS_CHECK(expr)                    \
  do {                                     \
    int _s = (expr);                       \
    if (_s != TCS_OK) return _s;           \
  } while (0)

// --------------------------------------------------------------------------------------
// tensor geometry: activations are NHWC.  "padded" tensors carry a 1-pixel circular halo:
// [B, H+2, W+2, C]; padded (py,px) holds image ((py-1) mod H, (px-1) mod W).
// --------------------------------------------------------------------------------------
enum Epilogue : int {
  EPI_RAW_STATS = 0,  // fp32 [B,H,W,N] + bias, plus GroupNorm partial sums
  EPI_PADDED = 1,     // T [B,H+2,W+2,ldo] + bias (+ residual), halo written
  EPI_PLAIN = 2,      // T [M, ldo] + bias
  EPI_GN_FUSED = 3,   // tcgen05 engine only: bias + GroupNorm + SiLU from TMEM -> padded bf16 (halo written)
  EPI_EPS = 4         // tcgen05 engine only: the 96 -> 1 output conv (N padded to 16) + CFG combine -> fp32 [n,4096]
};

struct ConvGeom {
  int B;          // images in this pass
  int H, W;       // OUTPUT spatial size
  int ksize;      // 1, 3 or 4
  int stride;     // 1 or 2
  int nsrc;       // 1 or 2 (skip-concat as split-K)
  int csrc[2];    // channels per source (multiples of 32)
  int in_pad[2];  // 1 = source is a padded tensor, 0 = plain
  int ntot;       // total output channels
  int split3 = 0;  // fp32 mode on the tensor pipe: operands as bf16 (hi, lo) pairs, a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi
                   // accumulated in fp32 (the dropped a_lo*w_lo term and the representation error are ~2^-17 relative)
  int ups = 0;     // tcgen05 engine, 3x3 stride-1 layers with a padded bf16 output: the source is the HALF-resolution padded
                   // tensor [B, H/2+2, W/2+2, C] and the bilinear x2 upsample (edge clamp) is blended on the way into shared memory
  int kx_in_n = 0; // 96 -> 1 output conv only: the three kx taps are three GEMM columns sharing ONE (unshifted) window per
                   // channel block; the epilogue sums column kx of pixel x + kx - 1 (circular in x)
};

struct EpiArgs {
  const float* bias;      // [ntot]
  void* out;              // see Epilogue
  float* partials;        // EPI_RAW_STATS: [B][slots][8][2]; EPI_GN_FUSED: the (value, flag) exchange words, 16 x 8 B per M tile
  const void* residual;   // EPI_PADDED optional: padded T [B,H+2,W+2,ntot]
  int ldo;                // channel pitch of `out`
  int slots;              // partial slots per image
  const float* gamma;     // EPI_GN_FUSED: GroupNorm affine [ntot]
  const float* beta;
  int* overflow;          // EPI_GN_FUSED: device flag, set when a pre-norm value may not fit the fp16 stash (|v| > 65504)
};

// device buffer that only ever grows (cudaMalloc / cudaFree outside the hot loop)
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  uint64_t* gen = nullptr;   // optional generation counter of the owner, bumped whenever the buffer moves
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  int ensure(size_t b) {
    if (b <= bytes) return TCS_OK;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    TCS_CUDA(cudaMalloc(&p, b));
    bytes = b;
    if (gen) ++*gen;
    return TCS_OK;
  }
  template <typename U> U* as() const { return static_cast<U*>(p); }
};

struct HostTensor { std::vector<float> v; std::vector<int64_t> shape; };

// GroupNorm partial-sum slots per image written by each conv engine
inline int tc_slots(int H, int W) { return (H * W / 128) * 4; }
inline int simt_slots(int H, int W) { return H * W / 64; }

}  // namespace tcs
