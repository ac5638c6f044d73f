// attn_tc.cu — SelfAttention2d (sde_score_model.py:136-167) of a 16x16x192 image as ONE tcgen05 kernel:
//   GroupNorm(8 groups) -> qkv 1x1 conv -> per head softmax(q k^T / sqrt(48)) v -> proj 1x1 conv -> + x
// q, k, v, the scores and the attention output never leave the SM (shared / tensor memory).  It replaces four
// launches (gn_image16, conv_tc<attn.qkv>, attention_mma_kernel on mma.sync, conv_tc<attn.proj>).
//
// Work split: an image (256 tokens) is a CLUSTER OF TWO CTAs, each owning 128 tokens = the 128 TMEM lanes = one M tile.
// Every CTA projects q, k, v of its own tokens; the k and v rows are written into BOTH CTAs' shared memory (DSMEM
// stores), so that each CTA holds all 256 keys of the current head.  All MMAs are cta_group::1 (M = 128):
//   qkv_h [128 x 144] = Xn [128 x 192] . W_h^T        A, B shared memory (K-major SWIZZLE_128B), per head h
//   S     [128 x 256] = Q'_h [128 x 48] . K_h^T       A = Q' in TENSOR MEMORY (bf16 packed, tcgen05.st), Q' = q log2e/sqrt(48)
//   O_h   [128 x  48] = P [128 x 256] . V_h           A = P in tensor memory, written IN PLACE over the S columns;
//                                                     B = V rows (keys) x 48 -> MN-major descriptor
//   Y     [128 x 192] = O [128 x 192] . Wproj^T       A = normalised O of all heads (shared memory)
// TMEM columns: [0,256) S / P / Y, [256,400) qkv_h, [400,424) Q' (packed), [424,472) O_h.
// Shared memory (214 KB): Xn 48 K | weights of the head 54 K (bulk-copied, pre-swizzled images; Wproj reuses this and the
// dead Xn region) | K 32 K | V 32 K (128-byte rows, 48 of 64 elements used) | O 48 K.
// Threads: warp 0 = control (one lane issues the bulk copies and every MMA), warps 1-8 = workers: warp w owns TMEM lane
// quarter w % 4 (32 tokens) and column half (w - 1) / 4 of whatever is being read, so a softmax row is shared by two
// threads (max and sum are combined through shared memory).
// The building blocks (A operand in tensor memory, MN-major partial atom, DSMEM-written operands, 1-D bulk copies)
// were probed on B200 first: tools/umma_attn_probe.cu, profiles/r2_attn_probe.txt.
#include <cooperative_groups.h>

#include <cstring>
#include <mutex>

#include "attn_tc.cuh"
#include "tc_ptx.cuh"

namespace cg = cooperative_groups;

namespace tcs {
namespace {

constexpr int AT_THREADS = 288;
constexpr int AT_TOK = 128;
constexpr int AT_PIMG = 18 * 18 * 192;                 // elements of a padded image
constexpr uint32_t XN_OFF = 0, W_OFF = 49152, K_OFF = 104448, V_OFF = 137216, O_OFF = 169984, AT_SMEM = 219136;
constexpr uint32_t KBLK = 16384;                       // a 128-row x 64-channel K block of an A operand
constexpr uint32_t WH_KB = 144 * 128, WH_BYTES = 3 * WH_KB;       // per-head q|k|v weights
constexpr uint32_t WP_KB = 192 * 128, WP_BYTES = 3 * WP_KB;       // projection weights
constexpr uint32_t S_COL = 0, QKV_COL = 256, Q_COL = 400, O_COL = 424;
constexpr float QSCALE = 0.14433756729740643f * 1.4426950408889634f;   // log2(e) / sqrt(48)
static_assert(W_OFF % 1024 == 0 && K_OFF % 1024 == 0 && V_OFF % 1024 == 0 && O_OFF % 1024 == 0, "swizzle atoms");
static_assert(W_OFF + WH_BYTES == K_OFF && 2 * WP_KB <= WH_BYTES && WP_KB <= W_OFF, "weight regions");

__host__ __device__ inline uint32_t sw128(uint32_t row, uint32_t ch) { return row * 128 + ((ch ^ (row & 7)) << 4); }

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ int halo_wrap16(int v) { return v == 0 ? 16 : (v == 15 ? -16 : 0); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(AT_THREADS, 1) attn_block_tc_kernel(const AttnTcParams p) {
  extern __shared__ uint8_t at_raw[];
  __shared__ uint64_t bar_w, bar_qkv, bar_s, bar_o, bar_y;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_part[AT_THREADS][2];
  __shared__ float s_mine[16], s_peer[16];
  __shared__ float s_mean[8], s_rstd[8];
  __shared__ float s_max[2][AT_TOK], s_sum[2][AT_TOK];
  __shared__ float s_bqkv[576], s_bproj[192], s_gamma[192], s_beta[192];

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank(), peer = rank ^ 1u;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform role dispatch
  const bool ctrl = warp == 0;
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t sb = ptx::smem_u32(sm);
  uint8_t* sm_peer = cluster.map_shared_rank(sm, peer);
  float* peer_part = cluster.map_shared_rank(&s_peer[0], peer);
  const uint32_t b_w = ptx::smem_u32(&bar_w), b_qkv = ptx::smem_u32(&bar_qkv), b_s = ptx::smem_u32(&bar_s),
                 b_o = ptx::smem_u32(&bar_o), b_y = ptx::smem_u32(&bar_y);

  for (int i = tid; i < 576; i += AT_THREADS) s_bqkv[i] = p.bias_qkv[i];
  for (int i = tid; i < 192; i += AT_THREADS) { s_bproj[i] = p.bias_proj[i]; s_gamma[i] = p.gamma[i]; s_beta[i] = p.beta[i]; }
  if (tid == 0) {
    ptx::mbar_init(b_w, 1); ptx::mbar_init(b_qkv, 1); ptx::mbar_init(b_s, 1); ptx::mbar_init(b_o, 1); ptx::mbar_init(b_y, 1);
    ptx::fence_barrier_init();
  }
  if (ctrl) ptx::tmem_alloc_512(ptx::smem_u32(&tmem_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  cluster.sync();   // the peer is running: its shared memory may be written from here on

  const int ncl = gridDim.x >> 1, cl = blockIdx.x >> 1;
  uint32_t ph_w = 0, ph_qkv = 0, ph_s = 0, ph_o = 0, ph_y = 0;
  if (ctrl && lane == 0 && cl < p.B) {
    ptx::mbar_expect_tx(b_w, WH_BYTES);
    ptx::bulk_load_1d(sb + W_OFF, p.wpack, WH_BYTES, b_w);
  }
  // worker coordinates (meaningless for the control warp)
  const int wq = warp & 3, hf = (warp - 1) >> 2;
  const int row = wq * 32 + lane;                    // token inside this CTA = TMEM lane
  const int gt = static_cast<int>(rank) * AT_TOK + row;   // token inside the image = key index
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
  // phase-0 coordinates: thread = (16-byte channel chunk, token lane)
  const int c0 = tid % 24, tl = tid / 24;

  for (int img = cl; img < p.B; img += ncl) {
    // clock64 phase profile (tests / tuning only): control lane and the first worker lane of CTA 0, second image of its loop
    long long* prof = (p.dbg && blockIdx.x == 0 && img == cl + ncl && (tid == 0 || tid == 32))
                          ? reinterpret_cast<long long*>(p.dbg + ATTN_DBG_FLOATS) + (tid ? ATTN_PROF_SLOTS : 0) : nullptr;
    int prof_i = 0;
#define AT_PROF() { if (prof && prof_i < ATTN_PROF_SLOTS) prof[prof_i++] = clock64(); }
    AT_PROF();
    const __nv_bfloat16* xb = p.x + static_cast<size_t>(img) * AT_PIMG;
    __nv_bfloat16* ob = p.out + static_cast<size_t>(img) * AT_PIMG;
    float* dbg = (p.dbg && img == 0) ? p.dbg : nullptr;
    // ---- phase 0: GroupNorm of the image (statistics over both CTAs' halves) -> Xn, the A operand of the projections
    {
      uint4 v[11];
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const int tok = tl + 12 * k;
        if (tok < AT_TOK) {
          const int g_t = static_cast<int>(rank) * AT_TOK + tok, y = g_t >> 4, x = g_t & 15;
          v[k] = __ldg(reinterpret_cast<const uint4*>(xb + ((y + 1) * 18 + x + 1) * 192 + c0 * 8));
        }
      }
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        if (tl + 12 * k < AT_TOK) {
          float f[8];
          unpack8(v[k], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) { s += f[j]; q += f[j] * f[j]; }
        }
      }
      s_part[tid][0] = s; s_part[tid][1] = q;
      __syncthreads();
      if (tid < 16) {   // fixed-order combine: run-to-run bit-identical
        const int g = tid >> 1, which = tid & 1;
        float a = 0.f;
        for (int t2 = 0; t2 < 12; ++t2)
          for (int cc = 0; cc < 3; ++cc) a += s_part[t2 * 24 + g * 3 + cc][which];
        s_mine[tid] = a;
        peer_part[tid] = a;
      }
      cluster.sync();
      if (tid < 8) {
        const float sa = rank == 0 ? s_mine[2 * tid] : s_peer[2 * tid], sb2 = rank == 0 ? s_peer[2 * tid] : s_mine[2 * tid];
        const float qa = rank == 0 ? s_mine[2 * tid + 1] : s_peer[2 * tid + 1], qb = rank == 0 ? s_peer[2 * tid + 1] : s_mine[2 * tid + 1];
        const double inv = 1.0 / (256.0 * 24.0);
        const double mean = (static_cast<double>(sa) + static_cast<double>(sb2)) * inv;
        double var = (static_cast<double>(qa) + static_cast<double>(qb)) * inv - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        s_mean[tid] = static_cast<float>(mean);
        s_rstd[tid] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(GN_EPS)));
      }
      __syncthreads();
      const int g = c0 / 3;
      const float mean = s_mean[g], rstd = s_rstd[g];
      float gm[8], bt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { gm[j] = s_gamma[c0 * 8 + j] * rstd; bt[j] = s_beta[c0 * 8 + j] - mean * gm[j]; }
      // (x - mean) rstd gamma + beta == x (rstd gamma) + (beta - mean rstd gamma)
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const int tok = tl + 12 * k;
        if (tok < AT_TOK) {
          float f[8];
          unpack8(v[k], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], gm[j], bt[j]);
          const uint4 u = pack8(f);
          *reinterpret_cast<uint4*>(sm + XN_OFF + (c0 >> 3) * KBLK + sw128(tok, c0 & 7)) = u;
          if (dbg) {
            float r[8];
            unpack8(u, r);
            for (int j = 0; j < 8; ++j) dbg[ATTN_DBG_XN + (static_cast<size_t>(rank) * AT_TOK + tok) * 192 + c0 * 8 + j] = r[j];
          }
        }
      }
      ptx::fence_proxy_async_all();
      ptx::tc_fence_before();
      __syncthreads();
      AT_PROF();   // 1: phase 0 done
    }
    // ---- per head -------------------------------------------------------------------------------------------------
    for (int h = 0; h < N_HEADS; ++h) {
      // (a) q|k|v of the head for this CTA's tokens
      if (ctrl) {
        if (lane == 0) {
          ptx::mbar_wait(b_w, ph_w);
          AT_PROF();   // c: weights there
          ptx::tc_fence_after();
          const uint32_t idesc = make_idesc(128, 144);
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) {
            const uint64_t ad = make_desc_sw128(sb + XN_OFF + kb * KBLK), bd = make_desc_sw128(sb + W_OFF + kb * WH_KB);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ptx::umma_bf16(tmem + QKV_COL, ad + 2 * ks, bd + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
          }
          ptx::umma_commit(b_qkv);
          AT_PROF();   // c: qkv issued
          ptx::mbar_wait(b_qkv, ph_qkv);   // the weights (and, after the last head, Xn) are consumed: fetch the next set
          if (h < N_HEADS - 1) {
            ptx::mbar_expect_tx(b_w, WH_BYTES);
            ptx::bulk_load_1d(sb + W_OFF, p.wpack + static_cast<size_t>(h + 1) * WH_BYTES, WH_BYTES, b_w);
          } else {
            const uint8_t* wp = p.wpack + static_cast<size_t>(N_HEADS) * WH_BYTES;
            ptx::mbar_expect_tx(b_w, WP_BYTES);
            ptx::bulk_load_1d(sb + W_OFF, wp, 2 * WP_KB, b_w);
            ptx::bulk_load_1d(sb + XN_OFF, wp + 2 * WP_KB, WP_KB, b_w);
          }
          AT_PROF();   // c: qkv done, next weights requested
        }
        __syncwarp();
      } else {
        ptx::mbar_wait(b_qkv, ph_qkv);
        AT_PROF();   // w: qkv accumulator ready
        ptx::tc_fence_after();
        // columns [72 hf, 72 hf + 72): 8-column chunks cj = 9 hf + j; chunks 0-5 are q, 6-11 k, 12-17 v
        float f[72];
        const uint32_t src = lane_addr + QKV_COL + 72 * hf;
        ptx::tmem_ld32(src, f);
        ptx::tmem_ld32(src + 32, f + 32);
        ptx::tmem_ld8(src + 64, f + 64);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int cj = 9 * hf + j, kind = cj / 6, within = cj - kind * 6;
          const float* bsrc = s_bqkv + kind * 192 + h * 48 + within * 8;
          float g8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) g8[i] = f[8 * j + i] + bsrc[i];
          if (dbg) {
            for (int i = 0; i < 8; ++i) dbg[ATTN_DBG_QKV + static_cast<size_t>(gt) * 576 + kind * 192 + h * 48 + within * 8 + i] = g8[i];
          }
          if (kind == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) g8[i] *= QSCALE;
            const uint4 u = pack8(g8);
            ptx::tmem_st4(lane_addr + Q_COL + within * 4, reinterpret_cast<const uint32_t*>(&u));
          } else {
            const uint4 u = pack8(g8);
            const uint32_t off = (kind == 1 ? K_OFF : V_OFF) + sw128(gt, within);
            *reinterpret_cast<uint4*>(sm + off) = u;
            *reinterpret_cast<uint4*>(sm_peer + off) = u;
          }
        }
        ptx::tmem_st_wait();
        AT_PROF();   // w: q, k, v converted and stored
        ptx::fence_proxy_async_all();
        ptx::tc_fence_before();
        AT_PROF();   // w: fences
      }
      ph_w ^= 1; ph_qkv ^= 1;
      cluster.sync();   // all 256 key / value rows of the head are in both CTAs' shared memory
      AT_PROF();   // both: cluster sync A passed
      // (b) S = Q' K^T
      if (ctrl) {
        if (lane == 0) {
          ptx::fence_proxy_async_all();
          ptx::tc_fence_after();
          const uint32_t idesc = make_idesc(128, 256);
          const uint64_t kd = make_desc_sw128(sb + K_OFF);
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) ptx::umma_bf16_ts(tmem + S_COL, tmem + Q_COL + 8 * ks, kd + 2 * ks, idesc, ks ? 1u : 0u);
          ptx::umma_commit(b_s);
          AT_PROF();   // c: S issued
        }
        __syncwarp();
      } else {
        ptx::mbar_wait(b_s, ph_s);
        AT_PROF();   // w: S ready
        ptx::tc_fence_after();
        // softmax over keys [128 hf, 128 hf + 128) of row `row`; P (bf16) overwrites the first half of those columns
        const uint32_t sc = lane_addr + S_COL + 128 * hf;
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float sv[32];
          ptx::tmem_ld32(sc + 32 * i, sv);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, sv[j]);
        }
        s_max[hf][row] = m;
        AT_PROF();   // w: pass 1 (max) done
        worker_bar();
        AT_PROF();   // w: worker barrier
        m = fmaxf(s_max[0][row], s_max[1][row]);
        float l = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float sv[32];
          ptx::tmem_ld32(sc + 32 * i, sv);
          ptx::tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = ptx::ex2_approx(sv[2 * j] - m), p1 = ptx::ex2_approx(sv[2 * j + 1] - m);
            l += p0 + p1;
            pk[j] = pack2(p0, p1);
          }
          ptx::tmem_st16(sc + 16 * i, pk);
        }
        s_sum[hf][row] = l;
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        AT_PROF();   // w: pass 2 (exp, P stored) done
      }
      ph_s ^= 1;
      __syncthreads();
      AT_PROF();   // both: block sync after softmax
      // (c) O_h = P V_h
      if (ctrl) {
        if (lane == 0) {
          ptx::tc_fence_after();
          const uint32_t idesc = make_idesc(128, 48) | IDESC_B_MN_MAJOR;
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const uint32_t a_col = S_COL + (ks < 8 ? 8 * ks : 128 + 8 * (ks - 8));
            ptx::umma_bf16_ts(tmem + O_COL, tmem + a_col, make_desc_sw128_mn(sb + V_OFF + ks * 2048), idesc, ks ? 1u : 0u);
          }
          ptx::umma_commit(b_o);
          AT_PROF();   // c: PV issued
        }
        __syncwarp();
      } else {
        ptx::mbar_wait(b_o, ph_o);
        AT_PROF();   // w: O ready
        ptx::tc_fence_after();
        float o[24];
        ptx::tmem_ld16(lane_addr + O_COL + 24 * hf, o);
        ptx::tmem_ld8(lane_addr + O_COL + 24 * hf + 16, o + 16);
        ptx::tmem_ld_wait();
        const float lsum = s_sum[0][row] + s_sum[1][row];
        const float inv = 1.0f / lsum;
        if (dbg && hf == 0) dbg[ATTN_DBG_L + static_cast<size_t>(gt) * 4 + h] = lsum;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float g8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) g8[i] = o[8 * j + i] * inv;
          const int gc = h * 6 + 3 * hf + j;   // 16-byte chunk of the 192-channel row
          *reinterpret_cast<uint4*>(sm + O_OFF + (gc >> 3) * KBLK + sw128(row, gc & 7)) = pack8(g8);
          if (dbg) {
            for (int i = 0; i < 8; ++i) dbg[ATTN_DBG_Y + static_cast<size_t>(gt) * 192 + gc * 8 + i] = g8[i];
          }
        }
        ptx::fence_proxy_async_all();
        ptx::tc_fence_before();
        AT_PROF();   // w: O converted
      }
      ph_o ^= 1;
      cluster.sync();   // the peer's MMAs have read this head's K / V rows: the next head may overwrite them
      AT_PROF();   // both: cluster sync B passed
    }
    // ---- projection + residual ------------------------------------------------------------------------------------
    if (ctrl) {
      if (lane == 0) {
        ptx::mbar_wait(b_w, ph_w);
        ptx::fence_proxy_async_all();
        ptx::tc_fence_after();
        const uint32_t idesc = make_idesc(128, 192);
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const uint64_t ad = make_desc_sw128(sb + O_OFF + kb * KBLK);
          const uint64_t bd = make_desc_sw128(kb < 2 ? sb + W_OFF + kb * WP_KB : sb + XN_OFF);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_bf16(tmem + S_COL, ad + 2 * ks, bd + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
        }
        ptx::umma_commit(b_y);
        AT_PROF();   // c: projection issued
        ptx::mbar_wait(b_y, ph_y);
        if (img + ncl < p.B) {   // first head's weights of the next image
          ptx::mbar_expect_tx(b_w, WH_BYTES);
          ptx::bulk_load_1d(sb + W_OFF, p.wpack, WH_BYTES, b_w);
        }
      }
      __syncwarp();
    } else {
      ptx::mbar_wait(b_y, ph_y);
      AT_PROF();   // w: Y ready
      ptx::tc_fence_after();
      const int y = gt >> 4, x = gt & 15;
      const size_t pix = static_cast<size_t>((y + 1) * 18 + x + 1) * 192 + 96 * hf;
      const long long wy = static_cast<long long>(halo_wrap16(y)) * 18 * 192, wx = static_cast<long long>(halo_wrap16(x)) * 192;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float yv[32];
        ptx::tmem_ld32(lane_addr + S_COL + 96 * hf + 32 * i, yv);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ch = 32 * i + 8 * j;
          float r[8], g8[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(xb + pix + ch)), r);
#pragma unroll
          for (int e = 0; e < 8; ++e) g8[e] = yv[8 * j + e] + s_bproj[96 * hf + ch + e] + r[e];
          const uint4 u = pack8(g8);
          __nv_bfloat16* dst = ob + pix + ch;
          *reinterpret_cast<uint4*>(dst) = u;
          if (wy) *reinterpret_cast<uint4*>(dst + wy) = u;
          if (wx) *reinterpret_cast<uint4*>(dst + wx) = u;
          if (wy && wx) *reinterpret_cast<uint4*>(dst + wy + wx) = u;
        }
      }
      ptx::tc_fence_before();
    }
    AT_PROF();   // both: epilogue / projection done
    ph_w ^= 1; ph_y ^= 1;
    __syncthreads();   // Y has been read and the projection weights consumed: the next image may overwrite Xn / S
  }
#undef AT_PROF
  ptx::tc_fence_before();
  cluster.sync();      // no DSMEM store may target a CTA that has exited
  if (ctrl) ptx::tmem_dealloc_512(tmem);
}

}  // namespace

size_t attn_tc_wpack_bytes() { return static_cast<size_t>(N_HEADS) * WH_BYTES + WP_BYTES; }

void attn_tc_pack_weights(const float* qkv_w, const float* proj_w, uint8_t* out) {
  std::memset(out, 0, attn_tc_wpack_bytes());
  auto put = [](uint8_t* base, uint32_t row, int ch, float v) {   // element (row, channel ch of 64) of a K block
    const __nv_bfloat16 b = __float2bfloat16(v);
    std::memcpy(base + sw128(row, ch >> 3) + (ch & 7) * 2, &b, 2);
  };
  for (int h = 0; h < N_HEADS; ++h)
    for (int r = 0; r < 144; ++r) {
      const int kind = r / 48, d = r % 48, orow = kind * 192 + h * 48 + d;   // torch.chunk(qkv, 3) then heads of 48
      for (int ci = 0; ci < 192; ++ci)
        put(out + static_cast<size_t>(h) * WH_BYTES + (ci >> 6) * WH_KB, r, ci & 63, qkv_w[orow * 192 + ci]);
    }
  uint8_t* wp = out + static_cast<size_t>(N_HEADS) * WH_BYTES;
  for (int r = 0; r < 192; ++r)
    for (int ci = 0; ci < 192; ++ci) put(wp + (ci >> 6) * WP_KB, r, ci & 63, proj_w[r * 192 + ci]);
}

int launch_attn_block_tc(const AttnTcParams& p, int sm_count, cudaStream_t st) {
  if (p.B <= 0) return TCS_OK;
  constexpr int smem = AT_SMEM + 1024;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int max_clusters = 0;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(attn_block_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (attr_err != cudaSuccess) return;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (sm_count / 2)); cfg.blockDim = dim3(AT_THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, attn_block_tc_kernel, &cfg) == cudaSuccess && n > 0) max_clusters = n;
    else cudaGetLastError();
  });
  TCS_CUDA(attr_err);
  int ncl = sm_count / 2;
  if (max_clusters > 0 && max_clusters < ncl) ncl = max_clusters;
  if (ncl > p.B) ncl = p.B;
  if (ncl < 1) ncl = 1;
  attn_block_tc_kernel<<<2 * ncl, AT_THREADS, smem, st>>>(p);
  TCS_CUDA(cudaGetLastError());
  return TCS_OK;
}

}  // namespace tcs
