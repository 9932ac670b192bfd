// multilevel_bwd.cu -- backward pass of the multi-level pooled sparse attention (second half).
//
// Reference: the Triton backward kernels of kernels/block_sparse_attn_kernel_with_backward_9_10.py (K9:695-1237, launch
// K9:1375-1576) behind `sparse_attention_fn` (K9:1578-1611).  They compute the plain gradient of the forward
// (tests/test_oracle_multilevel.py: the reference kernels == autograd through the oracle to 2e-5), including the path
// through the mean-pooled K/V copies.  With P = exp2(S * scale_log2 + log2(L) - lse2), Delta = rowsum(dO o O):
//     dP = dO V_t^T          dS = P o (dP - Delta) * softmax_scale
//     dQ += dS K_t           dK_t += dS^T Q          dV_t += P^T dO
// where K_t / V_t are the rows of ONE 128-key tensor-core tile (a level-1 block, or 2 / 4 / 8 pooled blocks of one
// level), exactly the tiles the forward assembles.
//
//   delta_kernel            Delta[b,h,r] = sum_d dO * O (fp32)
//   ml_bwd_kernel           one CTA per (query tile, head, batch): for every tile of the row's list the five GEMMs run on
//                           tcgen05 (operands in shared memory: Q, dO, K_t, V_t by TMA; P and dS written by the threads
//                           in the same 128-byte-swizzled layout; S / dP / dQ / dK_t / dV_t in TMEM), dK_t / dV_t are
//                           added to per-level fp32 accumulators in global memory with atomics (rows of different
//                           query tiles meet there).  First correct version: one tile in flight, no software
//                           pipeline -- the forward is the hot path, this is the training-side completion of row f4.
//   ml_bwd_combine_kernel   dK = dK_1 + unpool(dK_2)/2 + unpool(dK_4)/4 + unpool(dK_8)/8 (the gradient of the pair-mean
//                           pyramid, K9:1252-1270; gradients of the replicate-padded rows fold onto the last real row),
//                           rounded to the tensor dtype.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace blade {

constexpr int kBT = 128;  // tile edge (query rows, keys)

struct BwdMaps {
  CUtensorMap q, dO, k, v, kp[3], vp[3];
};

struct BwdParams {
  const int32_t* idx;
  const int4* cnt4;
  int64_t idx_stride;
  const float* lse;    // natural log, [B,H,S]
  const float* delta;  // [B,H,S]
  uint16_t* dq;
  int64_t dq_sb, dq_sh, dq_ss;
  float* dk_acc[4];  // level 1, 2, 4, 8: fp32 [B*H, rows[l], D]
  float* dv_acc[4];
  int64_t rows[4];
  int B, H, S, Sk, nq, nk;
  float scale_log2, scale;
};

__device__ __forceinline__ int bwd_tiles(const int4 c) { return c.x + ((c.y + 1) >> 1) + ((c.z + 3) >> 2) + ((c.w + 7) >> 3); }
__device__ __forceinline__ void bwd_tile(const int4 c, int j, int& lc, int& e0, int& ne) {
  const int t2 = (c.y + 1) >> 1, t4 = (c.z + 3) >> 2;
  if (j < c.x) { lc = 0; e0 = j; ne = 1; return; }
  j -= c.x;
  if (j < t2) { lc = 1; e0 = c.x + 2 * j; ne = min(2, c.y - 2 * j); return; }
  j -= t2;
  if (j < t4) { lc = 2; e0 = c.x + c.y + 4 * j; ne = min(4, c.z - 4 * j); return; }
  j -= t4;
  lc = 3; e0 = c.x + c.y + c.z + 8 * j; ne = min(8, c.w - 8 * j);
}

__device__ __forceinline__ void red_add_v4(float* addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
               "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}

template <bool IS_BF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  return IS_BF16 ? pack_bf16x2(lo, hi) : pack_f16x2(lo, hi);
}

// grid (ceil(rows/8)), 256 threads: one warp per (b,h,row)
template <bool IS_BF16>
__global__ void __launch_bounds__(256) delta_kernel(const uint16_t* __restrict__ o, const uint16_t* __restrict__ d_o,
                                                    int64_t o_sb, int64_t o_sh, int64_t o_ss, int64_t g_sb, int64_t g_sh,
                                                    int64_t g_ss, int H, int S, int D, int64_t total,
                                                    float* __restrict__ delta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= total) return;
  const int64_t bh = row / S, r = row % S;
  const int64_t b = bh / H, h = bh % H;
  const uint16_t* po = o + b * o_sb + h * o_sh + r * o_ss;
  const uint16_t* pg = d_o + b * g_sb + h * g_sh + r * g_ss;
  float acc = 0.f;
  for (int c = lane * 2; c < D; c += 64) {
    const uint32_t a = *reinterpret_cast<const uint32_t*>(po + c), g = *reinterpret_cast<const uint32_t*>(pg + c);
    float a0, a1, g0, g1;
    if (IS_BF16) {
      a0 = __uint_as_float(a << 16); a1 = __uint_as_float(a & 0xFFFF0000u);
      g0 = __uint_as_float(g << 16); g1 = __uint_as_float(g & 0xFFFF0000u);
    } else {
      const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&a)), fg = __half22float2(*reinterpret_cast<const __half2*>(&g));
      a0 = fa.x; a1 = fa.y; g0 = fg.x; g1 = fg.y;
    }
    acc = fmaf(a0, g0, acc);
    acc = fmaf(a1, g1, acc);
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) delta[row] = acc;
}

template <int D>
struct BwdSmem {
  static constexpr int kTile = kBT * D * 2;        // Q, dO, K_t, V_t tiles
  static constexpr int kPTile = kBT * kBT * 2;     // P, dS tiles (128 x 128)
  static constexpr int kQ = 0, kdO = kTile, kK = 2 * kTile, kV = 3 * kTile, kP = 4 * kTile, kdS = 4 * kTile + kPTile;
  static constexpr int kMisc = kdS + kPTile;
  static constexpr int kTotal = kMisc + 256 + 1024;
};

template <int D, bool IS_BF16>
__global__ void __launch_bounds__(128, 1) ml_bwd_kernel(const __grid_constant__ BwdMaps tm, const BwdParams p) {
  using L = BwdSmem<D>;
  constexpr int kSub = D / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sQ = smem + L::kQ, *sdO = smem + L::kdO, *sK = smem + L::kK, *sV = smem + L::kV, *sP = smem + L::kP,
          *sdS = smem + L::kdS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kMisc);   // [0] TMA, [1] MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int64_t bh = static_cast<int64_t>(b) * p.H + h;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t tS = tmem + lane_base, tdP = tmem + lane_base + kBT, tdQ = tmem + lane_base + 2 * kBT;

  const int64_t rowid = bh * p.nq + qb;
  const int4 c4 = __ldg(p.cnt4 + rowid);
  const int ntiles = bwd_tiles(c4);
  const int32_t* list = p.idx + rowid * p.idx_stride;
  const int r_glob = qb * kBT + tid;
  const bool valid_row = r_glob < p.S;
  const float lse2 = valid_row ? __ldg(p.lse + bh * p.S + r_glob) * 1.4426950408889634f : 0.f;
  const float dlt = valid_row ? __ldg(p.delta + bh * p.S + r_glob) : 0.f;

  uint32_t ph_ld = 0, ph_mma = 0;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 2 * L::kTile);
    for (int dh = 0; dh < kSub; ++dh) {
      tma_load_4d(sQ + dh * (kBT * 128), &tm.q, &bars[0], dh * 64, qb * kBT, h, b, kEvictFirst);
      tma_load_4d(sdO + dh * (kBT * 128), &tm.dO, &bars[0], dh * 64, qb * kBT, h, b, kEvictFirst);
    }
  }
  mbar_wait(&bars[0], ph_ld);
  ph_ld ^= 1;

  constexpr uint32_t idesc_s = make_idesc_f16(kBT, kBT, IS_BF16, false, false);   // S = Q K^T, dP = dO V^T
  constexpr uint32_t idesc_t = make_idesc_f16(kBT, D, IS_BF16, true, true);       // dV = P^T dO, dK = dS^T Q
  constexpr uint32_t idesc_q = make_idesc_f16(kBT, D, IS_BF16, false, true);      // dQ += dS K
  const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP),
                 adS = smem_u32(sdS);

  for (int j = 0; j < ntiles; ++j) {
    int lc, e0, ne;
    bwd_tile(c4, j, lc, e0, ne);
    const int rows_per = kBT >> lc;
    const int valid = ne * rows_per;
    const float bias = static_cast<float>(lc);
    if (tid == 0) {
      mbar_arrive_expect_tx(&bars[0], 2 * L::kTile);
      for (int g = 0; g < (1 << lc); ++g) {
        const int e = e0 + (g < ne ? g : ne - 1);
        const int kb = __ldg(list + e) & 0x0FFFFFFF;
        const CUtensorMap* mk = lc == 0 ? &tm.k : &tm.kp[lc - 1];
        const CUtensorMap* mv = lc == 0 ? &tm.v : &tm.vp[lc - 1];
        for (int dh = 0; dh < kSub; ++dh) {
          tma_load_4d(sK + dh * (kBT * 128) + g * rows_per * 128, mk, &bars[0], dh * 64, kb * rows_per, h, b, kEvictLast);
          tma_load_4d(sV + dh * (kBT * 128) + g * rows_per * 128, mv, &bars[0], dh * 64, kb * rows_per, h, b, kEvictLast);
        }
      }
    }
    mbar_wait(&bars[0], ph_ld);
    ph_ld ^= 1;
    tc_fence_after();
    if (tid == 0) {
      const uint64_t dq_ = make_smem_desc(aQ, 16, 1024, 2), dk_ = make_smem_desc(aK, 16, 1024, 2);
      const uint64_t ddo = make_smem_desc(adO, 16, 1024, 2), dv_ = make_smem_desc(aV, 16, 1024, 2);
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t koff = static_cast<uint32_t>((k >> 2) * (kBT * 128 / 16) + (k & 3) * 2);
        umma_ss(tmem, dq_ + koff, dk_ + koff, idesc_s, k > 0);
      }
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t koff = static_cast<uint32_t>((k >> 2) * (kBT * 128 / 16) + (k & 3) * 2);
        umma_ss(tmem + kBT, ddo + koff, dv_ + koff, idesc_s, k > 0);
      }
      tc_commit(&bars[1]);
    }
    mbar_wait(&bars[1], ph_mma);
    ph_mma ^= 1;
    tc_fence_after();

    // ---- P and dS rows (thread = query row), written as K-major SWIZZLE_128B tiles [128 rows][128 keys]
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t s[32], dp[32];
      tmem_ld32(tS + c * 32, s);
      tmem_ld32(tdP + c * 32, dp);
      tmem_wait_ld();
      uint32_t pw[16], dw[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float pv[2], dsv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int col = c * 32 + 2 * i + u;
          float pe = 0.f;
          if (valid_row && col < valid) pe = ex2_approx(fmaf(__uint_as_float(s[2 * i + u]), p.scale_log2, bias - lse2));
          pv[u] = pe;
          dsv[u] = pe * (__uint_as_float(dp[2 * i + u]) - dlt) * p.scale;
        }
        pw[i] = pack2<IS_BF16>(pv[0], pv[1]);
        dw[i] = pack2<IS_BF16>(dsv[0], dsv[1]);
      }
      // chunk c = keys [32c, 32c+32): sub-tile c/2, 16-byte chunks (c%2)*4 .. +3 of the row, XOR-swizzled with row%8
      const int sub = c >> 1;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ch = ((c & 1) * 4 + u) ^ (tid & 7);
        const int off = sub * (kBT * 128) + tid * 128 + ch * 16;
        *reinterpret_cast<uint4*>(sP + off) = make_uint4(pw[4 * u], pw[4 * u + 1], pw[4 * u + 2], pw[4 * u + 3]);
        *reinterpret_cast<uint4*>(sdS + off) = make_uint4(dw[4 * u], dw[4 * u + 1], dw[4 * u + 2], dw[4 * u + 3]);
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // A = P / dS as MN-major operands (M = key, contiguous in memory; K = query row): LBO = sub-tile stride, SBO = 8 rows
      const uint64_t aPt = make_smem_desc(aP, kBT * 128, 1024, 2), adSt = make_smem_desc(adS, kBT * 128, 1024, 2);
      const uint64_t bdO = make_smem_desc(adO, kBT * 128, 1024, 2), bQ = make_smem_desc(aQ, kBT * 128, 1024, 2);
      const uint64_t bK = make_smem_desc(aK, kBT * 128, 1024, 2);
      const uint64_t adSk = make_smem_desc(adS, 16, 1024, 2);   // dS as K-major A (K = key)
#pragma unroll
      for (int k = 0; k < kBT / 16; ++k) umma_ss(tmem, aPt + k * 128, bdO + k * 128, idesc_t, k > 0);            // dV_t
#pragma unroll
      for (int k = 0; k < kBT / 16; ++k) umma_ss(tmem + kBT, adSt + k * 128, bQ + k * 128, idesc_t, k > 0);      // dK_t
#pragma unroll
      for (int k = 0; k < kBT / 16; ++k) {
        const uint32_t koff = static_cast<uint32_t>((k >> 2) * (kBT * 128 / 16) + (k & 3) * 2);
        umma_ss(tmem + 2 * kBT, adSk + koff, bK + k * 128, idesc_q, (j > 0) || (k > 0));                          // dQ
      }
      tc_commit(&bars[1]);
    }
    mbar_wait(&bars[1], ph_mma);
    ph_mma ^= 1;
    tc_fence_after();

    // ---- dV_t / dK_t rows (thread = key of the tile) -> per-level fp32 accumulators
    {
      const int g = tid / rows_per, within = tid - g * rows_per;
      bool ok = tid < valid;
      int64_t prow = 0;
      if (ok) {
        const int kb = __ldg(list + e0 + g) & 0x0FFFFFFF;
        prow = static_cast<int64_t>(kb) * rows_per + within;
        ok = prow < p.rows[lc];   // level 1: zero-filled keys beyond the sequence carry no gradient
      }
      float* dst_v = p.dv_acc[lc] + (bh * p.rows[lc] + prow) * D;
      float* dst_k = p.dk_acc[lc] + (bh * p.rows[lc] + prow) * D;
#pragma unroll
      for (int cc = 0; cc < D / 32; ++cc) {
        uint32_t a[32], kk[32];
        tmem_ld32(tS + cc * 32, a);      // dV_t lives where S was
        tmem_ld32(tdP + cc * 32, kk);    // dK_t where dP was
        tmem_wait_ld();
        if (ok) {  // 16-byte vector reductions (REDG.E.ADD.F32x4): a quarter of the L2 atomic operations of scalar adds
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            red_add_v4(dst_v + cc * 32 + i, a[i], a[i + 1], a[i + 2], a[i + 3]);
            red_add_v4(dst_k + cc * 32 + i, kk[i], kk[i + 1], kk[i + 2], kk[i + 3]);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  // ---- dQ (a row without any tile gets zeros)
  {
    uint16_t* dst = p.dq + b * p.dq_sb + h * p.dq_sh + static_cast<int64_t>(r_glob) * p.dq_ss;
#pragma unroll
    for (int cc = 0; cc < D / 32; ++cc) {
      uint32_t a[32];
      if (ntiles > 0) {
        tmem_ld32(tdQ + cc * 32, a);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = 0u;
      }
      if (valid_row) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 w;
          w.x = pack2<IS_BF16>(__uint_as_float(a[8 * u + 0]), __uint_as_float(a[8 * u + 1]));
          w.y = pack2<IS_BF16>(__uint_as_float(a[8 * u + 2]), __uint_as_float(a[8 * u + 3]));
          w.z = pack2<IS_BF16>(__uint_as_float(a[8 * u + 4]), __uint_as_float(a[8 * u + 5]));
          w.w = pack2<IS_BF16>(__uint_as_float(a[8 * u + 6]), __uint_as_float(a[8 * u + 7]));
          reinterpret_cast<uint4*>(dst)[cc * 4 + u] = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// grid (ceil(Sk*D/8 / 256), B*H): thread = (row, 8 channels)
template <bool IS_BF16>
__global__ void __launch_bounds__(256) ml_bwd_combine_kernel(const float* __restrict__ a1, const float* __restrict__ a2,
                                                             const float* __restrict__ a4, const float* __restrict__ a8,
                                                             int64_t r1, int64_t r2, int64_t r4, int64_t r8, int Sk,
                                                             int D, int H, uint16_t* __restrict__ out, int64_t o_sb,
                                                             int64_t o_sh, int64_t o_ss) {
  const int64_t bh = blockIdx.y;
  const int lpr = D / 8;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t r = e / lpr;
  const int ch = static_cast<int>(e % lpr);
  if (r >= Sk) return;
  float acc[8];
  const float* p1 = a1 + (bh * r1 + r) * D + ch * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = p1[i];
  auto add_pooled = [&](int64_t rp) {
    const float* p2 = a2 + (bh * r2 + (rp >> 1)) * D + ch * 8;
    const float* p4 = a4 + (bh * r4 + (rp >> 2)) * D + ch * 8;
    const float* p8 = a8 + (bh * r8 + (rp >> 3)) * D + ch * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += p2[i] * 0.5f + p4[i] * 0.25f + p8[i] * 0.125f;
  };
  add_pooled(r);
  if (r == Sk - 1)                                    // replicate padding: the padded rows are copies of the last token
    for (int64_t rp = Sk; rp < r2 * 2; ++rp) add_pooled(rp);
  uint16_t* dst = out + (bh / H) * o_sb + (bh % H) * o_sh + r * o_ss + ch * 8;
  uint4 w;
  w.x = pack2<IS_BF16>(acc[0], acc[1]);
  w.y = pack2<IS_BF16>(acc[2], acc[3]);
  w.z = pack2<IS_BF16>(acc[4], acc[5]);
  w.w = pack2<IS_BF16>(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(dst) = w;
}

struct BwdWs {
  size_t delta, acc[2][4], total;
  int64_t rows[4];
};
static BwdWs bwd_carve(int64_t B, int64_t H, int64_t S, int64_t Sk, int64_t D) {
  BwdWs w{};
  const int64_t nk = ceil_div(Sk, 128);
  w.rows[0] = Sk;
  w.rows[1] = nk * 64;
  w.rows[2] = nk * 32;
  w.rows[3] = nk * 16;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  w.delta = take(B * H * S * 4);
  for (int t = 0; t < 2; ++t)
    for (int l = 0; l < 4; ++l) w.acc[t][l] = take(B * H * w.rows[l] * D * 4);
  w.total = off;
  return w;
}

}  // namespace blade

using namespace blade;

extern "C" size_t blade_multilevel_bwd_workspace_bytes(int64_t B, int64_t H, int64_t S, int64_t Sk, int64_t D) {
  return bwd_carve(B, H, S, Sk, D).total;
}

extern "C" int blade_multilevel_attn_bwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                                         const BladeTensor* k2, const BladeTensor* v2, const BladeTensor* k4,
                                         const BladeTensor* v4, const BladeTensor* k8, const BladeTensor* v8,
                                         const int32_t* idx, const int32_t* cnt4, int64_t idx_stride,
                                         const BladeTensor* out, const BladeTensor* d_out, const float* lse,
                                         float softmax_scale, BladeTensor* dq, BladeTensor* dk, BladeTensor* dv,
                                         void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (const BladeTensor* t : {q, k, v, out, d_out, static_cast<const BladeTensor*>(dq), static_cast<const BladeTensor*>(dk),
                               static_cast<const BladeTensor*>(dv)})
    if (int e = check_tensor16(t, "multilevel bwd tensor")) return e;
  BLADE_REQUIRE(idx && cnt4 && lse, BLADE_ERR_ARG, "idx / cnt4 / lse null");
  BLADE_REQUIRE((reinterpret_cast<uintptr_t>(cnt4) & 15) == 0, BLADE_ERR_ALIGN, "cnt4 not 16B aligned");
  const int64_t B = q->shape[0], H = q->shape[1], S = q->shape[2], D = q->shape[3], Sk = k->shape[2];
  for (int i = 0; i < 4; ++i) {
    BLADE_REQUIRE(out->shape[i] == q->shape[i] && d_out->shape[i] == q->shape[i] && dq->shape[i] == q->shape[i],
                  BLADE_ERR_SHAPE, "out / d_out / dq must have q's shape");
    BLADE_REQUIRE(v->shape[i] == k->shape[i] && dk->shape[i] == k->shape[i] && dv->shape[i] == k->shape[i],
                  BLADE_ERR_SHAPE, "v / dk / dv must have k's shape");
  }
  const bool bf = q->dtype == BLADE_BF16;
  const int nq = (int)ceil_div(S, kBT), nk = (int)ceil_div(Sk, kBT);
  const BwdWs w = bwd_carve(B, H, S, Sk, D);
  BLADE_REQUIRE(workspace && ws_bytes >= w.total, BLADE_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
                ws_bytes);
  BLADE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, BLADE_ERR_ALIGN, "workspace must be 1 KiB aligned");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  BLADE_CUDA_OK(cudaMemsetAsync(ws + w.acc[0][0], 0, w.total - w.acc[0][0], stream));

  BwdMaps tm;
  if (int e = make_tmap(&tm.q, q)) return e;
  if (int e = make_tmap(&tm.dO, d_out)) return e;
  if (int e = make_tmap(&tm.k, k)) return e;
  if (int e = make_tmap(&tm.v, v)) return e;
  const BladeTensor* pk[3] = {k2, k4, k8};
  const BladeTensor* pv[3] = {v2, v4, v8};
  for (int l = 0; l < 3; ++l) {
    const int rows = kBT >> (l + 1);
    for (const BladeTensor* t : {pk[l], pv[l]}) {
      if (int e = check_tensor16(t, "pyramid level")) return e;
      BLADE_REQUIRE(t->shape[0] == B && t->shape[1] == H && t->shape[3] == D && t->shape[2] >= static_cast<int64_t>(nk) * rows,
                    BLADE_ERR_SHAPE, "pyramid level %d: need [B,H,>=%lld,D]", 2 << l, (long long)nk * rows);
    }
    if (int e = make_tmap(&tm.kp[l], pk[l], rows)) return e;
    if (int e = make_tmap(&tm.vp[l], pv[l], rows)) return e;
  }
  float* delta = reinterpret_cast<float*>(ws + w.delta);
  {
    const int64_t total = B * H * S;
    const unsigned grid = static_cast<unsigned>(ceil_div(total, 8));
    const uint16_t *po = static_cast<const uint16_t*>(out->ptr), *pg = static_cast<const uint16_t*>(d_out->ptr);
    if (bf)
      delta_kernel<true><<<grid, 256, 0, stream>>>(po, pg, out->stride[0], out->stride[1], out->stride[2], d_out->stride[0],
                                                   d_out->stride[1], d_out->stride[2], (int)H, (int)S, (int)D, total, delta);
    else
      delta_kernel<false><<<grid, 256, 0, stream>>>(po, pg, out->stride[0], out->stride[1], out->stride[2], d_out->stride[0],
                                                    d_out->stride[1], d_out->stride[2], (int)H, (int)S, (int)D, total, delta);
    BLADE_CUDA_OK(cudaGetLastError());
  }
  BwdParams p{};
  p.idx = idx;
  p.cnt4 = reinterpret_cast<const int4*>(cnt4);
  p.idx_stride = idx_stride;
  p.lse = lse;
  p.delta = delta;
  p.dq = static_cast<uint16_t*>(dq->ptr);
  p.dq_sb = dq->stride[0];
  p.dq_sh = dq->stride[1];
  p.dq_ss = dq->stride[2];
  for (int l = 0; l < 4; ++l) {
    p.dk_acc[l] = reinterpret_cast<float*>(ws + w.acc[0][l]);
    p.dv_acc[l] = reinterpret_cast<float*>(ws + w.acc[1][l]);
    p.rows[l] = w.rows[l];
  }
  p.B = (int)B; p.H = (int)H; p.S = (int)S; p.Sk = (int)Sk; p.nq = nq; p.nk = nk;
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  dim3 grid(nq, static_cast<unsigned>(H), static_cast<unsigned>(B));
#define LAUNCH_BWD(DD, BF)                                                                                          \
  do {                                                                                                              \
    auto kern = ml_bwd_kernel<DD, BF>;                                                                              \
    BLADE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<DD>::kTotal));    \
    kern<<<grid, 128, BwdSmem<DD>::kTotal, stream>>>(tm, p);                                                         \
  } while (0)
  if (D == 128) { if (bf) LAUNCH_BWD(128, true); else LAUNCH_BWD(128, false); }
  else          { if (bf) LAUNCH_BWD(64, true); else LAUNCH_BWD(64, false); }
#undef LAUNCH_BWD
  BLADE_CUDA_OK(cudaGetLastError());
  {
    dim3 g2(static_cast<unsigned>(ceil_div(Sk * (D / 8), 256)), static_cast<unsigned>(B * H));
    for (int t = 0; t < 2; ++t) {
      BladeTensor* o = t == 0 ? dk : dv;
      float* const* acc = t == 0 ? p.dk_acc : p.dv_acc;
      if (bf)
        ml_bwd_combine_kernel<true><<<g2, 256, 0, stream>>>(acc[0], acc[1], acc[2], acc[3], w.rows[0], w.rows[1], w.rows[2],
                                                            w.rows[3], (int)Sk, (int)D, (int)H, static_cast<uint16_t*>(o->ptr),
                                                            o->stride[0], o->stride[1], o->stride[2]);
      else
        ml_bwd_combine_kernel<false><<<g2, 256, 0, stream>>>(acc[0], acc[1], acc[2], acc[3], w.rows[0], w.rows[1], w.rows[2],
                                                             w.rows[3], (int)Sk, (int)D, (int)H, static_cast<uint16_t*>(o->ptr),
                                                             o->stride[0], o->stride[1], o->stride[2]);
    }
    BLADE_CUDA_OK(cudaGetLastError());
  }
  return BLADE_OK;
}
