// attn_kernel.cu -- north-star kernel (b): block-sparse FlashAttention forward for sm_100a.
//
// Replaces the reference's two calls into the external mit-han-lab `block_sparse_attn_func`
// (sparse branch W:301-305 via W:278-309, pooled "global" branch W:21-24,348) and the LSE merge
// (W:351-370) with ONE persistent, warp-specialised kernel:
//
//   * work item  = one head x one PAIR of adjacent 128-row query tiles; the two tiles are two independent
//     "streams" that ping-pong on the tensor core (while one stream is in softmax the other is in MMA);
//   * per stream the KV sequence is [pooled tiles ..., selected blocks ...]: the pooled branch and the
//     sparse branch run back to back through the same pipeline with separate (m, l, O) state, because the
//     reference merges them with a chain of bf16 element-wise ops that a single fused softmax does not
//     reproduce; the pooled result is parked (bf16) in a small L2-resident workspace;
//   * warp 8 lane 0: TMA producer (Q tiles + a 5-deep ring of 128x128 K / V tiles, SWIZZLE_128B);
//     warp 9 lane 0: tcgen05.mma issuer -- S = Q K^T (SS, both operands K-major in smem, fp32 in TMEM),
//                    O += P V (TS: P read from TMEM where the softmax wrote it as bf16 over S, V MN-major);
//     warps 0-3 / 4-7: softmax warpgroups of stream 0 / 1: one query row per thread, tcgen05.ld of the
//                    128-column S row, online softmax in registers with lazy O rescale (threshold 2^8),
//                    tcgen05.st of P, final normalise + merge + (inverse-Gilbert) row store;
//   * TMEM map (512 columns): S0/P0 [0,128)  S1/P1 [128,256)  O0 [256,256+D)  O1 [256+D,256+2D).
//   * scheduling: items are claimed from a global atomic counter by the producer warp, which publishes
//     (item, block counts) through a 4-deep shared-memory queue to the issuer and the softmax warps -- rows keep
//     different numbers of blocks on real inputs, and a static round-robin lost 10-13 % there
//     (tools/imbalance_probe.py).  Without a workspace the same queue is fed with the static sequence.
//   * tail round: when the last round of pair items would occupy at most half of the CTAs, those pairs are
//     issued as "solo" items instead -- ONE query tile per CTA whose KV sequence is split between the two
//     streams (stream 0: pooled tiles + first part of the list, stream 1: the rest); warpgroup 0 reads both
//     accumulators straight from TMEM (same lanes) and folds them with the usual (m, l) rescale.  Each stream
//     is bound by its own softmax -> MMA latency chain, so halving the chain halves the tail round.
//
// All synchronisation is mbarrier based (TMA complete_tx, tcgen05.commit, thread arrivals).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace blade {

#ifdef BLADE_TRACE
#define TRACE(role, ev, n)                                                                          \
  do {                                                                                              \
    if (blockIdx.x == 0 && lane == 0 && (n) < 256) p.trace[((role)*8 + (ev)) * 256 + (n)] = clock64(); \
  } while (0)
#else
#define TRACE(role, ev, n) \
  do {                     \
  } while (0)
#endif

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;
constexpr int kThreads = 384;
constexpr int kSoftmaxThreads = 128;
constexpr float kRescaleThresh = 8.0f;  // log2 units
constexpr int kRegsSoftmax = 224;      // setmaxnreg split: 8 softmax warps x 224 + 4 other warps x 56 = 64512
constexpr int kRegsOther = 56;
constexpr int kMaxListSmem = 128;       // per-stream block ids cached in smem by the producer (longer lists: __ldg)
constexpr float kLn2 = 0.69314718055994530942f;
#ifndef BLADE_KV_HINT
#define BLADE_KV_HINT kEvictLast  // K/V tiles are re-read by every query tile of the head
#endif
#ifndef BLADE_SOLO_NUM
#define BLADE_SOLO_NUM 1        // dynamic queue: the last G * NUM / DEN pairs run as solo tiles
#define BLADE_SOLO_DEN 2
#endif
// of every 8 (score pair) exponentials, this many run as a Cody-Waite polynomial on the FMA pipe instead of MUFU.EX2
// (tools/microbench/softmax_loop.cu: 1 of 8 shortens the softmax loop by 7-8 %, 2 of 8 is break-even, more loses)
#ifndef BLADE_POLY_PAIRS
#define BLADE_POLY_PAIRS 0
#endif
constexpr int kPolyPairs = BLADE_POLY_PAIRS;
constexpr int kItemSlots = 4;          // depth of the per-CTA item queue
constexpr int kSchedBytes = 2048;      // head of the attention workspace: the global item counter + the arrival counters
                                       // of the query tiles whose KV list is split across two CTAs (ints 16 ..)
constexpr int kMaxSplitTiles = 496;    // (kSchedBytes / 4 - 16)
constexpr int kSplitDone = 1 << 16;    // arrival counter of a split tile: low bits = halves that claimed, this bit = the first half's dump is complete
template <int D>
constexpr size_t kSplitSlotBytes = static_cast<size_t>(D) * 768 + 2048;  // per split tile, see the softmax warps

template <int D>
struct SmemLayout {
  static constexpr int kTileBytes = kBlockN * D * 2;  // one K or V tile == one Q tile
  static constexpr int kStages = D == 128 ? 5 : 10;
  static constexpr int kQOff = 0;
  static constexpr int kKVOff = 2 * kTileBytes;
  static constexpr int kMiscOff = kKVOff + kStages * kTileBytes;
  static constexpr int kMiscBytes = 2048;
  static constexpr int kTotal = kMiscOff + kMiscBytes + 1024;  // + worst-case alignment slack
};

struct Misc {
  uint64_t q_full[2], q_empty[2];
  uint64_t s_full[2], p_lo[2], p_hi[2], o_full[2];  // P is handed over in two 64-key halves
  uint64_t kv_full[10], kv_empty[10];
  uint32_t tmem_base;
  uint32_t pad;
  uint16_t list[2][kMaxListSmem];
  float2 ml[kBlockM];  // solo items: stream 1's (m, l) per query row, handed to warpgroup 0
  // work queue: the producer warp claims items (atomic counter) and publishes (item, cnt0, cnt1) here
  int4 items[kItemSlots];
  uint64_t item_full[kItemSlots], item_empty[kItemSlots];
  int split_old;  // cross-CTA split: the arrival order of this half (broadcast from one thread to its warpgroup)
};
static_assert(sizeof(Misc) <= 2048, "misc smem region overflow");

struct AttnParams {
  const int32_t* idx;
  const int32_t* cnt;
  int64_t idx_stride;
  uint16_t* out;
  int64_t out_sb, out_sh, out_ss;
  float* lse;              // optional fp32 [B,H,S] (sparse-branch lse), may be null
  const int32_t* dst_row;  // optional
  uint4* park;             // [grid][2][D/8][128] 16-byte chunks (pooled-branch output)
  int B, H, S, Sk, nq, nk;
  int n_pool, n_pool_tiles;
  int num_items, pairs_per_head;
  int num_pair_items;      // items [0, num_pair_items) are tile pairs, the rest are solo tiles (two per pair id)
  int num_solo_pairs;      // solo tiles: two items per pair id
  int num_split_pairs;     // the LAST claims: tiles whose KV list is split across two CTAs (four items per pair id)
  uint8_t* split_ws;       // per split tile: parked pooled tile, two fp32 partial accumulators, (m, l), lse2
  int* split_cnt;          // arrival counters of the split tiles (zeroed with the item counter)
  int* sched;              // global item counter (zeroed before the launch); null = static round-robin
  float scale_log2;        // softmax_scale * log2(e)
  float log_gap_r;         // round_t(log(round_t(gap)))            (W:353-354)
  float gap;               // float(sample_gap) for the non-emulated merge
  int exact_merge;
  int sub64;  // 1: list entries carry a 4-bit quadrant mask in bits 28..31 (64x64 mask granularity on 128x128 tiles)
  // Ulysses push (BladePeers): output row `dst` (a token index) goes to peer dst / out_peer_rows, row dst % out_peer_rows
  // of that peer's [rows, H_total, D] buffer (pointers pre-offset to my first head; out_sh / out_ss describe that layout)
  uint16_t* out_peer[BLADE_MAX_PEERS];
  int out_peer_rows;  // 0 = off
  // MULTI (multi-level pooled attention, cogvideo_newattn.py N:154-207 + kernel K9:339-692): per (head, query tile) the
  // number of list entries of level 1, 2, 4, 8; the row's list is sorted by level, then block id
  const int4* cnt4;
  long long* trace;  // BLADE_TRACE builds: clock64 stamps of CTA 0 (tools/trace_attn.py)
};

template <bool IS_BF16>
__device__ __forceinline__ float round_t(float x) {
  return IS_BF16 ? __bfloat162float(__float2bfloat16_rn(x)) : __half2float(__float2half_rn(x));
}
template <bool IS_BF16>
__device__ __forceinline__ uint32_t pack_t(float lo, float hi) {
  return IS_BF16 ? pack_bf16x2(lo, hi) : pack_f16x2(lo, hi);
}
template <bool IS_BF16>
__device__ __forceinline__ uint32_t mul_t(uint32_t a, uint32_t b) {
  return IS_BF16 ? mul_bf16x2(a, b) : mul_f16x2(a, b);
}
template <bool IS_BF16>
__device__ __forceinline__ uint32_t add_t(uint32_t a, uint32_t b) {
  return IS_BF16 ? add_bf16x2(a, b) : add_f16x2(a, b);
}
template <bool IS_BF16>
__device__ __forceinline__ float2 unpack_t(uint32_t w) {
  if (IS_BF16) return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}

// K-major SWIZZLE_128B operand tile [128 rows][D] stored as D/64 sub-tiles of [128][64] (16 KB each):
// descriptor start offset (in 16-byte units) of the k-th 16-element K slice.
template <int D>
__device__ __forceinline__ uint32_t kmajor_koff(int k) {
  return static_cast<uint32_t>((k >> 2) * (kBlockN * 128 / 16) + (k & 3) * 2);
}

__device__ __forceinline__ uint4 shfl_xor_u4(const uint4& v, int mask) {
  return make_uint4(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask),
                    __shfl_xor_sync(0xffffffffu, v.z, mask), __shfl_xor_sync(0xffffffffu, v.w, mask));
}
// 4x4 transpose of 16-byte elements inside every group of 4 consecutive lanes.  In: a[u] = part u of this lane's
// row.  Out: a[j] = part (lane & 3) of the row of lane (lane & ~3) + j.  Two butterfly rounds, 16 SHFL.
// The epilogue uses it so that a warp-wide 16-byte store covers 8 rows x 64 contiguous bytes instead of 32 rows x
// 16 bytes: the row-per-lane store pattern cost ~2 000 L2 requests (5-9 k cycles) per tile.
__device__ __forceinline__ void transpose4x4(uint4 (&a)[4], int lane) {
  const bool hi = lane & 2, lo = lane & 1;
  // round 1 (xor 2): afterwards the lane holds parts {2hi, 2hi+1} of rows {me & ~2, me | 2}
  const uint4 s0 = hi ? a[0] : a[2], s1 = hi ? a[1] : a[3];
  const uint4 r0 = shfl_xor_u4(s0, 2), r1 = shfl_xor_u4(s1, 2);
  // t[row bit1][part bit0]
  const uint4 t00 = hi ? r0 : a[0], t01 = hi ? r1 : a[1];
  const uint4 t10 = hi ? a[2] : r0, t11 = hi ? a[3] : r1;
  // round 2 (xor 1): keep part 2hi+lo, trade the other part for the neighbour row's
  const uint4 q0 = lo ? t00 : t01, q1 = lo ? t10 : t11;
  const uint4 x0 = shfl_xor_u4(q0, 1), x1 = shfl_xor_u4(q1, 1);
  a[0] = lo ? x0 : t00;
  a[1] = lo ? t01 : x0;
  a[2] = lo ? x1 : t10;
  a[3] = lo ? t11 : x1;
}

// The item decode below is plain integer arithmetic shared by the three device roles AND by the host-side schedule
// checker (blade_debug_attn_schedule, tests/test_host_logic.py): host + device, read-only loads through ld_ro.
#define BLADE_HD __host__ __device__ __forceinline__
template <typename T>
BLADE_HD T ld_ro(const T* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
// One work item as the three roles see it: per stream t its query tile, how many pooled tiles it runs first,
// and which slice [off, off + ns) of the row's block list it owns.
struct Item {
  int bh;
  int qb[2], pt[2], off[2], ns[2];
  bool merge;  // solo item with work on both streams: warpgroup 0 folds stream 1's accumulator into its own
  bool split;  // one HALF of a query tile's KV sequence (the other half runs on another CTA); slot / half say which
  int slot, half;
};
// The per-row block counts are the only global-memory input of the decode; every role requests the NEXT item's
// counts at the top of the current item (item_counts) and turns them into an Item one iteration later (make_item),
// so no role stalls on a dependent load at an item boundary.
// Claim order -> (head, tile pair).  Inside a head the pairs are formed from the LAST query tile backwards -- pair 0 =
// tiles (nq-1, nq-2) -- so the rows the reference forces to full density (the last two query rows of every
// CogVideoX head, C:241-246: 139 blocks instead of ~16) share ONE item instead of idling the partner stream of two.
// Claims run: pair 0 of every head first (those items are up to 6x longer than the rest; started last they were the
// tail: list scheduling 724 steps vs 613 with them first, mean 608), then the remaining pairs head-major (K/V of
// ~2 heads stay L2-resident); the last num_solo_pairs claims are issued as two solo tiles each.
// ---- multi-level tiles: a level-L entry contributes the 128/L mean-pooled keys of its block; L consecutive entries of
// one level are packed into ONE 128-key tensor-core tile (the softmax is order independent), so a row runs
// n1 + ceil(n2/2) + ceil(n4/4) + ceil(n8/8) tiles.  Tile j -> level code lc = log2(L), first entry e0, entries ne.
BLADE_HD int multi_tiles(const int4 c) {
  return c.x + ((c.y + 1) >> 1) + ((c.z + 3) >> 2) + ((c.w + 7) >> 3);
}
__device__ __forceinline__ void multi_tile(const int4 c, int j, int& lc, int& e0, int& ne) {
  const int t2 = (c.y + 1) >> 1, t4 = (c.z + 3) >> 2;
  if (j < c.x) { lc = 0; e0 = j; ne = 1; return; }
  j -= c.x;
  if (j < t2) { lc = 1; e0 = c.x + 2 * j; ne = min(2, c.y - 2 * j); return; }
  j -= t2;
  if (j < t4) { lc = 2; e0 = c.x + c.y + 4 * j; ne = min(4, c.z - 4 * j); return; }
  j -= t4;
  lc = 3; e0 = c.x + c.y + c.z + 8 * j; ne = min(8, c.w - 8 * j);
}
struct MultiMaps {
  CUtensorMap k[3], v[3];  // K / V mean-pooled by 2, 4, 8 (box rows 64, 32, 16)
};

// Item classes in claim order: [0, npi) tile pairs | [npi, npi + 2 nsolo) solo tiles | [.., + 4 nsplit) HALF tiles --
// a query tile whose KV sequence is split across two CTAs, so that the last claims of a launch are quarter-length and
// the tail of the persistent schedule is flat (3 heads on 148 CTAs = 2.59 rounds of pairs used to cost 3.1).
// kind: 0 pair, 1 solo, 2 half; c = claim index of the pair slot; sub = which tile / half of it.
BLADE_HD void classify(const AttnParams& p, int item, int& kind, int& c, int& sub) {
  const int s0 = p.num_pair_items, s1 = s0 + 2 * p.num_solo_pairs;
  if (item < s0) { kind = 0; c = item; sub = 0; }
  else if (item < s1) { kind = 1; c = s0 + ((item - s0) >> 1); sub = (item - s0) & 1; }
  else { kind = 2; c = s0 + p.num_solo_pairs + ((item - s1) >> 2); sub = (item - s1) & 3; }
}
BLADE_HD int pair_id_of(const AttnParams& p, int c) {
  const int bh_n = p.B * p.H, pph = p.pairs_per_head;
  if (pph == 1) return c;
  if (c < bh_n) return c * pph;
  const int j = c - bh_n;
  return (j / (pph - 1)) * pph + 1 + j % (pph - 1);
}
BLADE_HD int tile_of(const AttnParams& p, int pair, int t) { return p.nq - 1 - (2 * pair + t); }
BLADE_HD void item_counts(const AttnParams& p, int item, int& c0, int& c1) {
  c0 = c1 = 0;
  if (item >= p.num_items) return;
  int kind, c, sub;
  classify(p, item, kind, c, sub);
  const int pid = pair_id_of(p, c);
  const int bh = pid / p.pairs_per_head, pair = pid % p.pairs_per_head;
  if (kind == 0) {
    if (p.cnt4) {  // multi-level: tile counts from the per-level entry counts
      const int4* row4 = p.cnt4 + static_cast<int64_t>(bh) * p.nq;
      c0 = multi_tiles(ld_ro(row4 + tile_of(p, pair, 0)));
      if (2 * pair + 1 < p.nq) c1 = multi_tiles(ld_ro(row4 + tile_of(p, pair, 1)));
      return;
    }
    const int32_t* row = p.cnt + static_cast<int64_t>(bh) * p.nq;
    c0 = ld_ro(row + tile_of(p, pair, 0));
    if (2 * pair + 1 < p.nq) c1 = ld_ro(row + tile_of(p, pair, 1));
  } else {
    const int qb = tile_of(p, pair, kind == 1 ? sub : (sub >> 1));
    if (qb >= 0) c0 = c1 = ld_ro(p.cnt + static_cast<int64_t>(bh) * p.nq + qb);
  }
}
BLADE_HD Item make_item(const AttnParams& p, int item, int c0, int c1) {
  Item it;
  const int npt = p.n_pool_tiles;
  int kind, cl, sub;
  classify(p, item, kind, cl, sub);
  const int pid = pair_id_of(p, cl);
  it.bh = pid / p.pairs_per_head;
  const int pair = pid % p.pairs_per_head;
  it.split = false;
  it.slot = it.half = 0;
  if (kind == 0) {
    it.merge = false;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int qb = tile_of(p, pair, t);
      const bool valid = qb >= 0;
      it.qb[t] = valid ? qb : p.nq;  // p.nq = "no tile" for the roles' qb < nq checks
      it.pt[t] = valid ? npt : 0;
      it.off[t] = 0;
      it.ns[t] = valid ? (t ? c1 : c0) : 0;
    }
  } else {
    const int qbr = tile_of(p, pair, kind == 1 ? sub : (sub >> 1));
    // a list of fewer than 4 entries is not worth a second CTA: half 0 runs it as an ordinary solo tile, half 1 is empty
    const bool halves = kind == 2 && c0 >= 4;
    const bool valid = qbr >= 0 && !(kind == 2 && !halves && (sub & 1));
    const int qb = valid ? qbr : p.nq;
    const int c = valid ? c0 : 0;
    // my share of the tile: [lo, lo + n) of its list, and the pooled tiles if I own them
    int lo = 0, n = c, mypt = valid ? npt : 0;
    if (halves && valid) {
      it.split = true;
      it.half = sub & 1;
      it.slot = (item - p.num_pair_items - 2 * p.num_solo_pairs) >> 1;
      // half 0 = pooled tiles + entries [0, ca); half 1 = entries [ca, c); both halves keep >= 2 entries, so each has
      // work on both streams (it.merge) -- the epilogue's half protocol relies on that
      int ca = (c - npt) / 2;
      ca = ca < 2 ? 2 : (ca > c - 2 ? c - 2 : ca);
      if (it.half) { lo = ca; n = c - ca; mypt = 0; } else { n = ca; }
    }
    // two streams: balance pooled + a  against  n - a tiles
    int a = n;
    if (n >= 2) {
      a = (n - mypt) / 2;
      a = a < 1 ? 1 : (a > n - 1 ? n - 1 : a);
    }
    it.qb[0] = it.qb[1] = qb;
    it.pt[0] = mypt;
    it.pt[1] = 0;
    it.off[0] = lo;
    it.ns[0] = a;
    it.off[1] = lo + a;
    it.ns[1] = n - a;
    it.merge = n - a > 0;
  }
  return it;
}

// ================================================================================================
template <int D, bool IS_BF16, bool POOLED, bool MULTI>
__device__ __forceinline__ void attn_body(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                          const CUtensorMap& tmKp, const CUtensorMap& tmVp, const MultiMaps* mm,
                                          const AttnParams& p) {
  using L = SmemLayout<D>;
  constexpr int kStages = L::kStages;
  constexpr int kTileBytes = L::kTileBytes;
  constexpr int kSub = D / 64;  // 64-column sub-tiles per operand tile

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + L::kQOff;
  uint8_t* sKV = smem + L::kKVOff;
  Misc* mz = reinterpret_cast<Misc*>(smem + L::kMiscOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(&mz->q_full[t], 1);
      mbar_init(&mz->q_empty[t], 1);
      mbar_init(&mz->s_full[t], 1);
      mbar_init(&mz->p_lo[t], kSoftmaxThreads / 32);  // one arrival per softmax warp
      mbar_init(&mz->p_hi[t], kSoftmaxThreads / 32);
      mbar_init(&mz->o_full[t], 1);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&mz->kv_full[s], 1);
      mbar_init(&mz->kv_empty[s], 1);
    }
    for (int s = 0; s < kItemSlots; ++s) {
      mbar_init(&mz->item_full[s], 1);
      mbar_init(&mz->item_empty[s], 1 + 2 * kSoftmaxThreads / 32);  // issuer warp + every softmax warp
    }
    fence_barrier_init();
  }
  if (warp == 11) tmem_alloc<512>(&mz->tmem_base);
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    if (p.n_pool_tiles) {
      tma_prefetch_desc(&tmKp);
      tma_prefetch_desc(&tmVp);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = mz->tmem_base;

  const int npt = p.n_pool_tiles;

  // Register split (setmaxnreg) lives INSIDE each role branch: ptxas applies the lowest bound it sees to all
  // code after a control-flow merge.
  if (warp == 8) {
    // ============================== TMA producer (warp-uniform; one elected lane issues) ==============
    setmaxnreg_dec<kRegsOther>();
    uint32_t slot = 0, ph = 0;  // ring position / phase of the NEXT load
    uint32_t q_it[2] = {0, 0};
    int n_claimed = 0;
    auto claim = [&]() {
      int v = 0;
      if (lane == 0) v = p.sched ? atomicAdd(p.sched, 1) : static_cast<int>(blockIdx.x + n_claimed * gridDim.x);
      ++n_claimed;
      return __shfl_sync(0xffffffffu, v, 0);
    };
    int c0n, c1n;
    int next_item = claim();
    item_counts(p, next_item, c0n, c1n);
    for (int qk = 0;; ++qk) {
      const int item = next_item;
      const int qslot = qk & (kItemSlots - 1);
      mbar_wait(&mz->item_empty[qslot], ((qk / kItemSlots) & 1) ^ 1);
      if (lane == 0) {
        mz->items[qslot] = make_int4(item < p.num_items ? item : -1, c0n, c1n, 0);
        mbar_arrive(&mz->item_full[qslot]);
      }
      __syncwarp();
      if (item >= p.num_items) break;
      const Item it = make_item(p, item, c0n, c1n);
      next_item = claim();                          // one item ahead: the counts are in flight during this item
      item_counts(p, next_item, c0n, c1n);
      const int bh = it.bh;
      const int b = bh / p.H, h = bh % p.H;
      int ns[2], nt[2];
      const int32_t* lists[2];
      int4 c4[2] = {make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0)};
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int64_t row = static_cast<int64_t>(bh) * p.nq + it.qb[t];
        ns[t] = it.ns[t];
        nt[t] = it.pt[t] + it.ns[t];
        lists[t] = p.idx + row * p.idx_stride + it.off[t];
        int n_ent = ns[t];
        if (MULTI && ns[t] > 0) {
          c4[t] = __ldg(p.cnt4 + row);
          n_ent = c4[t].x + c4[t].y + c4[t].z + c4[t].w;  // list entries (>= tiles)
        }
        // cooperative, coalesced fetch of the block-id list into smem (private to this warp)
        for (int j = lane; j < n_ent && j < kMaxListSmem; j += 32)
          mz->list[t][j] = static_cast<uint16_t>(__ldg(lists[t] + j) & 0x0FFFFFFF);
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (nt[t] == 0) continue;
        mbar_wait(&mz->q_empty[t], (q_it[t] & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&mz->q_full[t], kTileBytes);
#pragma unroll
          for (int dh = 0; dh < kSub; ++dh)
            tma_load_4d(sQ + t * kTileBytes + dh * (kBlockM * 128), &tmQ, &mz->q_full[t], dh * 64,
                        it.qb[t] * kBlockM, h, b, kEvictFirst);
        }
        __syncwarp();
        ++q_it[t];
      }
      auto load_tile = [&](int t, int j, bool is_v) {
        mbar_wait(&mz->kv_empty[slot], ph ^ 1);
        const CUtensorMap* map;
        int row;
        if (MULTI) {
          // one 128-key tile = L sub-boxes of 128/L pooled rows, one per list entry of this level; a last, partly
          // filled tile repeats its final entry (finite data; the softmax masks those columns)
          int lc, e0, ne;
          multi_tile(c4[t], j, lc, e0, ne);
          const int rows_per = kBlockN >> lc;
          map = lc == 0 ? (is_v ? &tmV : &tmK) : (is_v ? &mm->v[lc - 1] : &mm->k[lc - 1]);
          if (elect_one()) {
            mbar_arrive_expect_tx(&mz->kv_full[slot], kTileBytes);
            for (int gsub = 0; gsub < (1 << lc); ++gsub) {
              const int e = e0 + (gsub < ne ? gsub : ne - 1);
              const int kb = e < kMaxListSmem ? static_cast<int>(mz->list[t][e]) : (__ldg(lists[t] + e) & 0x0FFFFFFF);
#pragma unroll
              for (int dh = 0; dh < kSub; ++dh)
                tma_load_4d(sKV + slot * kTileBytes + dh * (kBlockN * 128) + gsub * rows_per * 128, map,
                            &mz->kv_full[slot], dh * 64, kb * rows_per, h, b, BLADE_KV_HINT);
            }
          }
          __syncwarp();
          if (++slot == kStages) {
            slot = 0;
            ph ^= 1;
          }
          return;
        }
        if (j < it.pt[t]) {
          map = is_v ? &tmVp : &tmKp;
          row = j * kBlockN;
        } else {
          const int jj = j - it.pt[t];
          const int kb = jj < kMaxListSmem ? static_cast<int>(mz->list[t][jj]) : (__ldg(lists[t] + jj) & 0x0FFFFFFF);
          map = is_v ? &tmV : &tmK;
          row = kb * kBlockN;
        }
        if (elect_one()) {
#ifdef BLADE_DIAG_HALF_KV
          // DIAGNOSTIC build only (wrong results): fetch half of every K/V tile -- is the kernel bound by L2 -> SM bytes?
          mbar_arrive_expect_tx(&mz->kv_full[slot], kTileBytes / kSub);
          tma_load_4d(sKV + slot * kTileBytes, map, &mz->kv_full[slot], 0, row, h, b, BLADE_KV_HINT);
#else
          mbar_arrive_expect_tx(&mz->kv_full[slot], kTileBytes);
#pragma unroll
          for (int dh = 0; dh < kSub; ++dh)
            tma_load_4d(sKV + slot * kTileBytes + dh * (kBlockN * 128), map, &mz->kv_full[slot], dh * 64, row, h, b,
                        BLADE_KV_HINT);
#endif
        }
        __syncwarp();
        if (++slot == kStages) {
          slot = 0;
          ph ^= 1;
        }
      };
      const int nmax = nt[0] > nt[1] ? nt[0] : nt[1];
      if (nt[0]) load_tile(0, 0, false);
      if (nt[1]) load_tile(1, 0, false);
      for (int j = 0; j < nmax; ++j) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j < nt[t]) {
            load_tile(t, j, true);
            if (j + 1 < nt[t]) load_tile(t, j + 1, false);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ============================== MMA issuer (warp-uniform; one elected lane issues) ==============
    setmaxnreg_dec<kRegsOther>();
    constexpr uint32_t idesc_qk = make_idesc_f16(kBlockM, kBlockN, IS_BF16, false, false);
    constexpr uint32_t idesc_pv = make_idesc_f16(kBlockM, D, IS_BF16, false, true);
    uint32_t slot = 0, ph = 0;  // ring position / phase of the NEXT tile to consume
    uint32_t g[2] = {0, 0};
    uint32_t q_it[2] = {0, 0};
    const uint32_t sQ_addr = smem_u32(sQ), sKV_addr = smem_u32(sKV);
    for (int qk = 0;; ++qk) {
      const int qslot = qk & (kItemSlots - 1);
      mbar_wait(&mz->item_full[qslot], (qk / kItemSlots) & 1);
      const int4 qe = mz->items[qslot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&mz->item_empty[qslot]);
      if (qe.x < 0) break;
      const int item = qe.x;
      const Item it = make_item(p, item, qe.y, qe.z);
      const int nt[2] = {it.pt[0] + it.ns[0], it.pt[1] + it.ns[1]};
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (nt[t] == 0) continue;
        mbar_wait(&mz->q_full[t], q_it[t] & 1);
        ++q_it[t];
      }
      auto issue_qk = [&](int t, bool last) {
        mbar_wait(&mz->kv_full[slot], ph);
        tc_fence_after();
        TRACE(2 + t, 2, g[t]);
        if (elect_one()) {
          const uint64_t adesc = make_smem_desc(sQ_addr + t * kTileBytes, 16, 1024, 2);
          const uint64_t bdesc = make_smem_desc(sKV_addr + slot * kTileBytes, 16, 1024, 2);
          const uint32_t tS = tmem_base + t * kBlockN;
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_ss(tS, adesc + kmajor_koff<D>(k), bdesc + kmajor_koff<D>(k), idesc_qk, k > 0);
          tc_commit(&mz->kv_empty[slot]);
          tc_commit(&mz->s_full[t]);
          if (last) tc_commit(&mz->q_empty[t]);
        }
        __syncwarp();
        if (++slot == kStages) {
          slot = 0;
          ph ^= 1;
        }
      };
      auto issue_pv = [&](int t, bool fresh) {
        // split-P: the PV GEMM over keys 0..63 is issued as soon as the first half of P is in TMEM, while the
        // softmax warpgroup still exponentiates keys 64..127
        mbar_wait(&mz->p_lo[t], g[t] & 1);
        TRACE(2 + t, 0, g[t]);
        mbar_wait(&mz->kv_full[slot], ph);
        tc_fence_after();
        TRACE(2 + t, 1, g[t]);
        // V tile [kv 128][D] as D/64 sub-tiles [128][64]; MN-major B: LBO = sub-tile stride, SBO = 8 kv rows
        const uint64_t bdesc = make_smem_desc(sKV_addr + slot * kTileBytes, kBlockN * 128, 1024, 2);
        const uint32_t tP = tmem_base + t * kBlockN;
        const uint32_t tO = tmem_base + 2 * kBlockN + t * D;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBlockN / 32; ++k)
            umma_ts(tO, tP + k * 8, bdesc + k * (16 * 128 / 16), idesc_pv, (!fresh) || k > 0);
        }
        __syncwarp();
        mbar_wait(&mz->p_hi[t], g[t] & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = kBlockN / 32; k < kBlockN / 16; ++k)
            umma_ts(tO, tP + k * 8, bdesc + k * (16 * 128 / 16), idesc_pv, 1);
          tc_commit(&mz->kv_empty[slot]);
          tc_commit(&mz->o_full[t]);
        }
        __syncwarp();
        if (++slot == kStages) {
          slot = 0;
          ph ^= 1;
        }
        ++g[t];
      };
      const int nmax = nt[0] > nt[1] ? nt[0] : nt[1];
      if (nt[0]) issue_qk(0, nt[0] == 1);
      if (nt[1]) issue_qk(1, nt[1] == 1);
      for (int j = 0; j < nmax; ++j) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j < nt[t]) {
            issue_pv(t, j == 0 || j == it.pt[t]);
            if (j + 1 < nt[t]) issue_qk(t, j + 2 == nt[t]);
          }
        }
      }
    }
  } else if (warp < 8) {
    // ============================== softmax warpgroups ==============================
    setmaxnreg_inc<kRegsSoftmax>();
    const int t = warp >> 2;
    const int wq = warp & 3;
    const int row_in_tile = wq * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + t * kBlockN;
    const uint32_t tO = tmem_base + lane_base + 2 * kBlockN + t * D;
    uint4* park = p.park + (static_cast<size_t>(blockIdx.x) * 2 + t) * (D / 8) * kBlockM;
    uint32_t g = 0;
    const float sl2 = p.scale_log2;
    const int pool_tail = p.n_pool - (npt - 1) * kBlockN;
    const int seq_tail = p.Sk - (p.nk - 1) * kBlockN;  // valid keys in the last key block

    for (int qk = 0;; ++qk) {
      const int qslot = qk & (kItemSlots - 1);
      mbar_wait(&mz->item_full[qslot], (qk / kItemSlots) & 1);
      const int4 qe = mz->items[qslot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&mz->item_empty[qslot]);
      if (qe.x < 0) break;
      const int item = qe.x;
      const Item it = make_item(p, item, qe.y, qe.z);
      const int bh = it.bh;
      const int b = bh / p.H, h = bh % p.H;
      const int qb = t ? it.qb[1] : it.qb[0];  // selects, not indexing: keeps Item in registers
      if (qb >= p.nq) continue;
      const bool solo = item >= p.num_pair_items;
      if (solo && t == 1 && !it.merge) continue;  // nothing to split: stream 0 runs the whole tile alone
      const int my_pt = t ? it.pt[1] : it.pt[0];
      const int ns = t ? it.ns[1] : it.ns[0];
      const int32_t* my_list =
          p.idx + (static_cast<int64_t>(bh) * p.nq + qb) * p.idx_stride + (t ? it.off[1] : it.off[0]);
      int4 c4 = make_int4(0, 0, 0, 0);
      int last_kb = -1;
      if (MULTI) {
        if (ns > 0) c4 = __ldg(p.cnt4 + static_cast<int64_t>(bh) * p.nq + qb);
      } else {
        last_kb = ns > 0 ? (__ldg(my_list + ns - 1) & 0x0FFFFFFF) : -1;
      }
      // MULTI keeps the zero-filled keys beyond the sequence IN the softmax (score 0, value 0), like the reference
      // kernel's masked loads (K9:108-119,155-178); only partly filled pooled tiles are masked (per tile, below)
      const int sparse_tail = (!MULTI && last_kb == p.nk - 1) ? seq_tail : kBlockN;
      float lse2 = 0.f;
      // half tiles (KV sequence split across two CTAs) keep their shared state in a per-tile slot of the workspace:
      // [D/8][128] parked pooled tile | [D][128] fp32 partial of the half that finishes first | [128] (m, l) | [128] lse2
      // (pointers formed where they are used: nothing extra stays live across the tile loops)
      auto slot_ptr = [&]() { return p.split_ws + static_cast<size_t>(it.slot) * kSplitSlotBytes<D>; };
      auto park_ptr = [&]() { return it.split ? reinterpret_cast<uint4*>(slot_ptr()) : park; };
      if (wq == 0 && sparse_tail > 0) TRACE(t, 7, g);  // item decoded (both dependent loads done)

      for (int phase = (my_pt ? 0 : 1); phase < 2; ++phase) {
        const int ntile = phase == 0 ? my_pt : ns;
        const int tail_valid = phase == 0 ? pool_tail : sparse_tail;
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < ntile; ++j, ++g) {
          mbar_wait(&mz->s_full[t], g & 1);
          tc_fence_after();
          if (wq == 0) TRACE(t, 0, g);
          uint32_t s[4][32];
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, s[c]);
          tmem_wait_ld();
          if (wq == 0) TRACE(t, 1, g);
          int valid = (j == ntile - 1) ? tail_valid : kBlockN;
          float bias = 0.f;  // MULTI: + log2(L) on the scaled score (K9: `qk += log(level)`, here in log2 units)
          if (MULTI) {
            int lc, e0, ne;
            multi_tile(c4, j, lc, e0, ne);
            valid = ne * (kBlockN >> lc);
            bias = static_cast<float>(lc);
          }
          float mxc[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent max chains
          // 64x64 mask granularity (block_size 64): bits (2*rowhalf + colhalf) of the entry's quadrant mask
          unsigned cmask = 3u;
          if (p.sub64 && phase == 1)
            cmask = (static_cast<unsigned>(__ldg(my_list + j)) >> (28 + 2 * (wq >> 1))) & 3u;
          if (valid < kBlockN || cmask != 3u) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c * 32 + i >= valid || !((cmask >> (c >> 1)) & 1u)) s[c][i] = __float_as_uint(-INFINITY);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; i += 2)
              mxc[c] = fmaxf(mxc[c], fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
          const float mx = fmaxf(fmaxf(mxc[0], mxc[1]), fmaxf(mxc[2], mxc[3])) * sl2 + bias;
          if (j == 0) {
            m = mx;
          } else if (m == -INFINITY) {
            m = mx;  // every earlier tile was fully masked for this row half (sub64 lists): O and l are still 0
          } else {
            const float m_cand = fmaxf(m, mx);
            const bool need = (m_cand - m) > kRescaleThresh;
            if (__any_sync(0xffffffffu, need)) {
              const float m_new = need ? m_cand : m;
              const float alpha = ex2_approx(m - m_new);
              mbar_wait(&mz->o_full[t], (g - 1) & 1);  // PV of the previous tile has retired
              tc_fence_after();
#pragma unroll
              for (int c = 0; c < D / 32; ++c) {
                uint32_t o[32];
                tmem_ld32(tO + c * 32, o);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st32(tO + c * 32, o);
              }
              tmem_wait_st();
              l *= alpha;
              m = m_new;
            }
          }
          if (wq == 0) TRACE(t, 2, g);
          const float neg_m = ((m == -INFINITY) ? 0.f : -m) + bias;  // all-masked so far: exp2(-inf - 0) = 0, never NaN
          // polynomial exponentials only on tiles without masked (-inf) scores: the polynomial clamps instead of returning 0
          const bool unmasked = kPolyPairs > 0 && valid == kBlockN && cmask == 3u;
          // packed fp32x2 math: x = s * scale - m and the row-sum accumulation take one issue slot per PAIR
          const uint64_t sl2_2 = pack_f32x2(sl2, sl2), negm_2 = pack_f32x2(neg_m, neg_m);
          uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};  // four independent packed sum chains
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint64_t x = fma_f32x2(pack_u32x2(s[c][2 * i], s[c][2 * i + 1]), sl2_2, negm_2);
              float p0, p1;
              if (kPolyPairs > 0 && (i & 7) < kPolyPairs && unmasked) {
                ex2_poly_x2(x, p0, p1);   // this pair on the FMA pipe (MUFU is the softmax's critical resource)
              } else {
                p0 = ex2_approx(lo_f32(x));
                p1 = ex2_approx(hi_f32(x));
              }
              ls2[i & 3] = add_f32x2(ls2[i & 3], pack_f32x2(p0, p1));
              pk[i] = pack_t<IS_BF16>(p0, p1);
            }
            if (c == 2) {
              // split-P: keys 0..63 of P (chunks 0,1) were stored a chunk of exponentials ago, so this wait is
              // free; the tensor core starts PV on them while chunks 2,3 are still being produced
              tmem_wait_st();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&mz->p_lo[t]);
            }
            tmem_st16(tS + c * 16, pk);
          }
          const uint64_t lsa = add_f32x2(add_f32x2(ls2[0], ls2[1]), add_f32x2(ls2[2], ls2[3]));
          float ls[4] = {lo_f32(lsa), hi_f32(lsa), 0.f, 0.f};
          l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
          if (wq == 0) TRACE(t, 3, g);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&mz->p_hi[t]);
          if (wq == 0) TRACE(t, 4, g);
        }
        // ---- phase finalisation: the last PV of the phase has to retire first (o_full)
        if (it.merge && t == 1) {
          // solo item, second half of the list: hand (m, l) to warpgroup 0, which reads O1 from TMEM itself
          mbar_wait(&mz->o_full[t], (g - 1) & 1);
          mz->ml[row_in_tile] = make_float2(m, l);
          tc_fence_before();
          named_bar_arrive<1, 2 * kSoftmaxThreads>();
          named_bar_sync<2, 2 * kSoftmaxThreads>();  // O1 has been read: the next item may overwrite it
          continue;
        }
        if (phase == 0) {
          // pooled branch done: normalise, round to the tensor dtype and park it until the sparse branch is done
          mbar_wait(&mz->o_full[t], (g - 1) & 1);
          tc_fence_after();
          const float inv_l = 1.0f / l;
          lse2 = (m + log2f(l)) * kLn2;  // natural-log LSE of the scaled scores
          uint4* const park_cur = park_ptr();
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint4 w;
              w.x = pack_t<IS_BF16>(__uint_as_float(o[8 * u + 0]) * inv_l, __uint_as_float(o[8 * u + 1]) * inv_l);
              w.y = pack_t<IS_BF16>(__uint_as_float(o[8 * u + 2]) * inv_l, __uint_as_float(o[8 * u + 3]) * inv_l);
              w.z = pack_t<IS_BF16>(__uint_as_float(o[8 * u + 4]) * inv_l, __uint_as_float(o[8 * u + 5]) * inv_l);
              w.w = pack_t<IS_BF16>(__uint_as_float(o[8 * u + 6]) * inv_l, __uint_as_float(o[8 * u + 7]) * inv_l);
              park_cur[(c * 4 + u) * kBlockM + row_in_tile] = w;
            }
          }
        } else {
          // Everything the epilogue needs from global memory (the parked pooled-branch result, 16 bytes per 8
          // columns, and the inverse-permutation row) is requested BEFORE waiting for the last PV: one overlapped L2
          // round trip instead of a dependent one per 8 columns (which cost ~10 us per item).
          const int r = qb * kBlockM + row_in_tile;
          const bool store = r < p.S;
          // ---- half tile: this CTA ran ONE HALF of the query tile's KV sequence (always a solo item with work on both
          // streams).  One atomic per half decides the roles.  The half that finishes FIRST (role 1) writes its
          // normalised accumulator (both streams folded) and (m, l) into the tile's slot and raises the slot's flag; the
          // half that finishes SECOND (role 2) waits for that flag -- the partner is past its last MMA and only has its
          // dump left, a bounded few-microsecond wait -- folds the partner's partial in as a third stream and runs the
          // ordinary epilogue (pooled merge, inverse permutation, row store).
          int role = 0;
          float m2 = -INFINITY, l2 = 0.f, f0 = 0.f, f1 = 0.f;
          uint8_t* const sslot = it.split ? slot_ptr() : nullptr;
          const uint4* const park_cur = it.split ? reinterpret_cast<const uint4*>(sslot) : park;
          if (it.split) {
            mbar_wait(&mz->o_full[t], (g - 1) & 1);   // my own work is done before I claim (keeps role 2's wait short)
            if (row_in_tile == 0) mz->split_old = atomicAdd(p.split_cnt + it.slot, 1);
            named_bar_sync<3, kSoftmaxThreads>();
            role = *reinterpret_cast<volatile int*>(&mz->split_old) == 0 ? 1 : 2;
            if (role == 2) {
              if (lane == 0) {
                uint32_t polls = 0;
                while ((*reinterpret_cast<volatile int*>(p.split_cnt + it.slot) & kSplitDone) == 0) {
                  __nanosleep(64);
                  if (++polls > (1u << 24)) __trap();
                }
              }
              __syncwarp();
              __threadfence();
              const float2 mlp = __ldcg(reinterpret_cast<const float2*>(sslot + D * 768) + row_in_tile);
              m2 = mlp.x;
              l2 = mlp.y;
              if (POOLED && it.half == 1) lse2 = __ldcg(reinterpret_cast<const float*>(sslot + D * 768 + 1024) + row_in_tile);
            }
          }
          uint4 parked[D / 8];
          if (POOLED) {  // compile-time: a runtime predicate here sends the array through local memory
#pragma unroll
            for (int i = 0; i < D / 8; ++i) parked[i] = __ldcg(park_cur + i * kBlockM + row_in_tile);
          }
          int dst = r;
          if (store && p.dst_row) dst = __ldg(p.dst_row + r);
          // (m, l) are final once the last tile's exponentials are summed, i.e. BEFORE its PV retires: the merge
          // weights and the row pointers are computed under the o_full wait (solo items redo them after the fold)
          float w0, w1 = 0.f, lse, alpha = 1.f, oma = 0.f;
          auto merge_weights = [&]() {
            lse = (m + log2f(l)) * kLn2;  // natural-log LSE of the scaled scores
            if (POOLED) {
              if (p.exact_merge) {
                // W:351-370 op by op, every intermediate rounded to the tensor dtype (lse already is, W:309)
                const float a = round_t<IS_BF16>(lse);
                const float b2 = round_t<IS_BF16>(round_t<IS_BF16>(lse2) + p.log_gap_r);
                const float mxw = fmaxf(a, b2);
                const float e1 = round_t<IS_BF16>(expf(round_t<IS_BF16>(a - mxw)));
                const float e2 = round_t<IS_BF16>(expf(round_t<IS_BF16>(b2 - mxw)));
                alpha = round_t<IS_BF16>(e1 / round_t<IS_BF16>(e1 + e2));
                oma = round_t<IS_BF16>(1.0f - alpha);
              } else {
                alpha = 1.0f / (1.0f + p.gap * expf(lse2 - lse));
                oma = 1.0f - alpha;
              }
            }
          };
          auto weights = [&](float a0) {
            const float inv_l = 1.0f / l;
            w0 = a0 * inv_l;
            w1 *= inv_l;
            merge_weights();
          };
          if (!it.merge) weights(1.f);
          uint16_t* orow = nullptr;
          if (store) {
            if (p.out_peer_rows) {  // stores over NVLink, overlapped with the next tile's MMAs
              const int pp = dst / p.out_peer_rows;
              orow = p.out_peer[pp] + h * p.out_sh + static_cast<int64_t>(dst - pp * p.out_peer_rows) * p.out_ss;
            } else {
              orow = p.out + b * p.out_sb + h * p.out_sh + static_cast<int64_t>(dst) * p.out_ss;
            }
          }
          // output row pointers of the 4 lanes of my group (null = row beyond S): after the 4x4 transpose lane L
          // stores part L & 3 of each of them
          uint16_t* orow_j[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            orow_j[j] = reinterpret_cast<uint16_t*>(
                __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(orow), (lane & ~3) + j));
          mbar_wait(&mz->o_full[t], (g - 1) & 1);
          tc_fence_after();
          if (wq == 0) TRACE(t, 5, g - 1);
          uint32_t ob[2][32];  // the accumulator chunks are loaded one ahead of the arithmetic
          tmem_ld32(tO, ob[0]);
          if (it.merge) {
            named_bar_sync<1, 2 * kSoftmaxThreads>();
            tc_fence_after();
            const float2 ml1 = mz->ml[row_in_tile];
            const float mm = fmaxf(m, ml1.x);
            const float a0 = (m == -INFINITY) ? 0.f : ex2_approx(m - mm);
            w1 = (ml1.x == -INFINITY) ? 0.f : ex2_approx(ml1.x - mm);
            l = a0 * l + w1 * ml1.y;
            m = mm;
            weights(a0);
            if (role == 2) {
              // the two halves' normalised partials X0, X1 combine as X0 * f0 + X1 * f1 with everything evaluated in HALF
              // order (not arrival order), so the result does not depend on which CTA finished first
              const float mA = it.half ? m2 : m, lA = it.half ? l2 : l, mB = it.half ? m : m2, lB = it.half ? l : l2;
              const float mm2 = fmaxf(mA, mB);
              const float sA = (mA == -INFINITY) ? 0.f : ex2_approx(mA - mm2) * lA;
              const float sB = (mB == -INFINITY) ? 0.f : ex2_approx(mB - mm2) * lB;
              l = sA + sB;
              m = mm2;
              const float inv_l = 1.0f / l;
              f0 = sA * inv_l;
              f1 = sB * inv_l;
              merge_weights();
            }
          }
          const uint32_t alpha2 = pack_t<IS_BF16>(alpha, alpha), oma2 = pack_t<IS_BF16>(oma, oma);
          if (store && p.lse && role != 1) p.lse[static_cast<int64_t>(bh) * p.S + r] = lse;
          // role 1's dump / role 2's third stream: fp32 [D][128] after the slot's parked pooled tile
          float* const part = reinterpret_cast<float*>(sslot + D * 256) + row_in_tile;
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t(&o)[32] = ob[c & 1];
            if (it.merge) {
              uint32_t o1[32];
              tmem_ld32(tO + D + c * 32, o1);  // stream 1's accumulator: same lanes, next D columns
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i)
                o[i] = __float_as_uint(__uint_as_float(o[i]) * w0 + __uint_as_float(o1[i]) * w1);
            } else {
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * w0);
            }
            if (c + 1 < D / 32) tmem_ld32(tO + (c + 1) * 32, ob[(c + 1) & 1]);
            if (role == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) __stcg(part + (c * 32 + i) * kBlockM, __uint_as_float(o[i]));
              continue;
            }
            if (role == 2) {
              float o2[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) o2[i] = __ldcg(part + (c * 32 + i) * kBlockM);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x0 = it.half ? o2[i] : __uint_as_float(o[i]), x1 = it.half ? __uint_as_float(o[i]) : o2[i];
                o[i] = __float_as_uint(fmaf(x1, f1, x0 * f0));
              }
            }
            uint4 wq4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint32_t wv[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float f0 = __uint_as_float(o[8 * u + 2 * i]), f1 = __uint_as_float(o[8 * u + 2 * i + 1]);
                if (!POOLED) {
                  wv[i] = pack_t<IS_BF16>(f0, f1);
                } else {
                  const uint4 w2 = parked[c * 4 + u];
                  const uint32_t pw = i == 0 ? w2.x : (i == 1 ? w2.y : (i == 2 ? w2.z : w2.w));
                  if (p.exact_merge) {
                    // out1 * alpha + out2 * (1 - alpha) as three tensor-dtype ops (W:368-370): packed native
                    // 16-bit multiplies / add, each rounding once like the element-wise torch ops they mirror
                    const uint32_t x = mul_t<IS_BF16>(pack_t<IS_BF16>(f0, f1), alpha2);
                    const uint32_t y = mul_t<IS_BF16>(pw, oma2);
                    wv[i] = add_t<IS_BF16>(x, y);
                  } else {
                    const float2 o2 = unpack_t<IS_BF16>(pw);
                    wv[i] = pack_t<IS_BF16>(f0 * alpha + o2.x * oma, f1 * alpha + o2.y * oma);
                  }
                }
              }
              wq4[u] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            }
            transpose4x4(wq4, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (orow_j[j]) reinterpret_cast<uint4*>(orow_j[j])[c * 4 + (lane & 3)] = wq4[j];
            if (wq == 0 && c < 4) TRACE(2 + t, 3 + c, g - 1);
          }
          if (it.merge) {
            tc_fence_before();
            named_bar_arrive<2, 2 * kSoftmaxThreads>();
          }
          if (role == 1) {
            // publish: (m, l) of this half (+ the pooled branch's lse if it ran here), then the flag
            __stcg(reinterpret_cast<float2*>(sslot + D * 768) + row_in_tile, make_float2(m, l));
            if (POOLED && it.half == 0) __stcg(reinterpret_cast<float*>(sslot + D * 768 + 1024) + row_in_tile, lse2);
            __threadfence();
            named_bar_sync<3, kSoftmaxThreads>();
            if (row_in_tile == 0) atomicOr(p.split_cnt + it.slot, kSplitDone);
          }
          if (wq == 0) TRACE(t, 6, g - 1);
        }
        // O_t / S_t are handed back implicitly: the next PV of this stream waits for our next p_full.
        tc_fence_before();
      }
    }
  } else {
    setmaxnreg_dec<kRegsOther>();  // warps 10, 11: idle (warp 11 owns the TMEM allocation)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 11) tmem_dealloc<512>(tmem_base);
}

template <int D, bool IS_BF16, bool POOLED>
__global__ void __launch_bounds__(kThreads, 1)
asa_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmKp,
                const __grid_constant__ CUtensorMap tmVp, const AttnParams p) {
  attn_body<D, IS_BF16, POOLED, false>(tmQ, tmK, tmV, tmKp, tmVp, nullptr, p);
}

// Multi-level pooled sparse attention (reference: the Triton `_fwd_kernel`, K9:339-692, behind
// cogvideo_newattn.py N:210-267): the same pipeline, tiles assembled from the K/V pyramid, `+ log2(level)` in the
// softmax, one softmax per row (no pooled branch, no merge).
template <int D, bool IS_BF16>
__global__ void __launch_bounds__(kThreads, 1)
asa_multilevel_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ MultiMaps mm,
                           const AttnParams p) {
  attn_body<D, IS_BF16, false, true>(tmQ, tmK, tmV, tmK, tmV, &mm, p);
}

// ================================================================================================
// bring-up probes: one tile through the same descriptors / TMEM layouts as the main kernel
// ================================================================================================
template <int D>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float* p_in,
             float* out, int mode /*0 = QK^T, 1 = PV*/) {
  constexpr int kTileBytes = kBlockN * D * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kTileBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const int row = warp * 32 + lane;

  if (mode == 1) {
    // P (fp32 [128][128]) -> bf16 pairs in TMEM columns [0,64)
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[16];
      for (int i = 0; i < 16; ++i)
        pk[i] = pack_bf16x2(p_in[row * kBlockN + c * 32 + 2 * i], p_in[row * kBlockN + c * 32 + 2 * i + 1]);
      tmem_st16(tmem_base + lane_base + c * 16, pk);
    }
    tmem_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    mbar_arrive_expect_tx(&bars[0], mode == 0 ? 2 * kTileBytes : kTileBytes);
    for (int dh = 0; dh < D / 64; ++dh) {
      if (mode == 0) tma_load_4d(sA + dh * (kBlockM * 128), &tmA, &bars[0], dh * 64, 0, 0, 0, kEvictNormal);
      tma_load_4d(sB + dh * (kBlockN * 128), &tmB, &bars[0], dh * 64, 0, 0, 0, kEvictNormal);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    if (mode == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kBlockM, kBlockN, true, false, false);
      const uint64_t adesc = make_smem_desc(smem_u32(sA), 16, 1024, 2);
      const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 1024, 2);
      for (int k = 0; k < D / 16; ++k)
        umma_ss(tmem_base + 128, adesc + kmajor_koff<D>(k), bdesc + kmajor_koff<D>(k), idesc, k > 0);
    } else {
      constexpr uint32_t idesc = make_idesc_f16(kBlockM, D, true, false, true);
      const uint64_t bdesc = make_smem_desc(smem_u32(sB), kBlockN * 128, 1024, 2);
      for (int k = 0; k < kBlockN / 16; ++k)
        umma_ts(tmem_base + 128, tmem_base + k * 8, bdesc + k * (16 * 128 / 16), idesc, k > 0);
    }
    tc_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int ncols = mode == 0 ? kBlockN : D;
  for (int c = 0; c < ncols / 32; ++c) {
    uint32_t o[32];
    tmem_ld32(tmem_base + lane_base + 128 + c * 32, o);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) out[row * ncols + c * 32 + i] = __uint_as_float(o[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// ================================================================================================
// host side: tensor maps + launch
// ================================================================================================
#ifdef BLADE_TRACE
static long long* g_trace_buf = nullptr;
#endif
static thread_local int g_sched_prezeroed = 0;  // blade_asa_forward zeroes the item counter ahead of the mask kernels
void attn_sched_prezeroed() { g_sched_prezeroed = 1; }
static thread_local int g_sub64_next = 0;  // set by blade_block_sparse_attn64_fwd / blade_asa_attn64_fwd for one launch
void attn_next_sub64() { g_sub64_next = 1; }


int device_sm_count() {
  static int n[64] = {0};  // per device: one process may drive several GPUs
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = n[dev & 63];
  if (!c) cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
  return c;
}

// attention workspace: [0, kSchedBytes) item counter + arrival counters | per CTA and stream one parked pooled-branch tile |
// per half-split query tile (at most one per CTA) one slot of shared state
static size_t attn_park_region(int64_t D) { return static_cast<size_t>(device_sm_count()) * 2 * (D / 8) * kBlockM * 16; }
static size_t attn_split_slot(int64_t D) { return D == 128 ? kSplitSlotBytes<128> : kSplitSlotBytes<64>; }
static int attn_max_split_tiles() { return device_sm_count() < kMaxSplitTiles ? device_sm_count() : kMaxSplitTiles; }
size_t attn_park_bytes(int64_t D) {
  return kSchedBytes + attn_park_region(D) + static_cast<size_t>(attn_max_split_tiles()) * attn_split_slot(D);
}
size_t attn_sched_bytes() { return kSchedBytes; }

// Tail classes of a launch: how many of the np tile pairs are issued as solo tiles (two items per pair) and as half tiles
// (four items per pair) behind the ordinary pair items.  Pure host arithmetic (also used by the schedule checker).
static void plan_tail(int np, int G, bool no_split, bool can_x, bool multi, bool dynamic, int xmax, int env_xs, int env_solo,
                      int& split, int& xsplit) {
  split = xsplit = 0;
  if (no_split || multi) return;  // multi-level rows are not split across the two streams
  const int r = np % G;
  if (can_x) {
    if (np >= G) {
      // measured (tools/sweep_xsplit.sh, profiles/r02z_sweep_xsplit.log): G/8 pairs of half tiles behind G pairs of
      // solo tiles is best or within noise of best for 12 and 3 Wan heads, uniform and non-uniform rows, and CogVideoX
      const int want_x = env_xs >= 0 ? env_xs : G / 8, want_s = env_solo >= 0 ? env_solo : G;
      xsplit = want_x < xmax ? want_x : xmax;
      xsplit = xsplit < np ? xsplit : np;
      split = want_s < np - xsplit ? want_s : np - xsplit;
    } else if (2 * np <= G && 4 * np > G) {
      split = np;                                   // one round of solo tiles already fills the machine
    } else {
      xsplit = np < xmax ? np : xmax;               // few pairs: quarter items
      split = np - xsplit < G / 2 ? np - xsplit : G / 2;
    }
  } else if (np < G) split = 2 * np <= G ? np : 0;
  else if (dynamic) split = (G * BLADE_SOLO_NUM / BLADE_SOLO_DEN) < np ? G * BLADE_SOLO_NUM / BLADE_SOLO_DEN : np;
  else split = (r > 0 && 2 * r <= G) ? r : 0;
}

static float round_host(float x, bool bf16) {
  return bf16 ? __bfloat162float(__float2bfloat16_rn(x)) : __half2float(__float2half_rn(x));
}

int launch_attn(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v, const int32_t* idx,
                const int32_t* cnt, int64_t idx_stride, const BladeTensor* k_pool, const BladeTensor* v_pool,
                int32_t sample_gap, BladeTensor* out, float* lse, const int32_t* dst_row, float softmax_scale,
                int exact_merge, void* workspace, size_t ws_bytes, cudaStream_t stream, const BladePeers* peers,
                const MultiLevelArgs* multi) {
  // one-shot flags set by the callers for exactly this launch: consumed up front, so that an early error return
  // cannot leak them into the next call on this thread
  const bool sched_prezeroed = g_sched_prezeroed != 0;
  const int sub64 = g_sub64_next;
  g_sched_prezeroed = 0;
  g_sub64_next = 0;
  if (int e = check_tensor16(q, "q")) return e;
  if (int e = check_tensor16(k, "k")) return e;
  if (int e = check_tensor16(v, "v")) return e;
  if (int e = check_tensor16(out, "out")) return e;
  BLADE_REQUIRE(idx && (cnt || multi), BLADE_ERR_ARG, "idx/cnt null");
  const int64_t B = q->shape[0], H = q->shape[1], S = q->shape[2], D = q->shape[3];
  const int64_t Sk = k->shape[2];
  for (int i = 0; i < 4; ++i) {
    BLADE_REQUIRE(v->shape[i] == k->shape[i], BLADE_ERR_SHAPE, "k/v shapes differ in dim %d", i);
    BLADE_REQUIRE(out->shape[i] == q->shape[i], BLADE_ERR_SHAPE, "out/q shapes differ in dim %d", i);
    if (i != 2) BLADE_REQUIRE(k->shape[i] == q->shape[i], BLADE_ERR_SHAPE, "q/k shapes differ in dim %d", i);
  }
  BLADE_REQUIRE(q->dtype == k->dtype && q->dtype == v->dtype && q->dtype == out->dtype, BLADE_ERR_DTYPE,
                "q/k/v/out dtypes differ");
  BLADE_REQUIRE(S < (1 << 30) && Sk < (1 << 30), BLADE_ERR_SHAPE, "sequence too long");
  const int nq = (int)ceil_div(S, kBlockM), nk = (int)ceil_div(Sk, kBlockN);
  BLADE_REQUIRE(nk <= 65535, BLADE_ERR_SHAPE, "too many key blocks");
  BLADE_REQUIRE(idx_stride >= 1, BLADE_ERR_ARG, "idx_stride");
  const bool pooled = sample_gap > 0 && k_pool && v_pool && k_pool->ptr && v_pool->ptr;
  int n_pool = 0;
  if (pooled) {
    if (int e = check_tensor16(k_pool, "k_pool")) return e;
    if (int e = check_tensor16(v_pool, "v_pool")) return e;
    n_pool = (int)k_pool->shape[2];
    BLADE_REQUIRE(v_pool->shape[2] == n_pool && k_pool->shape[3] == D && k_pool->shape[0] == B &&
                      k_pool->shape[1] == H && n_pool >= 1,
                  BLADE_ERR_SHAPE, "pooled k/v shape mismatch");
    BLADE_REQUIRE(workspace && ws_bytes >= attn_park_bytes(D), BLADE_ERR_WORKSPACE,
                  "attention workspace too small: need %zu bytes", attn_park_bytes(D));
    BLADE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, BLADE_ERR_ALIGN, "workspace not 16B aligned");
  }
  const bool bf = q->dtype == BLADE_BF16;

  CUtensorMap tmQ, tmK, tmV, tmKp, tmVp;
  if (int e = make_tmap(&tmQ, q)) return e;
  if (int e = make_tmap(&tmK, k)) return e;
  if (int e = make_tmap(&tmV, v)) return e;
  if (pooled) {
    if (int e = make_tmap(&tmKp, k_pool)) return e;
    if (int e = make_tmap(&tmVp, v_pool)) return e;
  } else {
    tmKp = tmK;
    tmVp = tmV;
  }

  AttnParams p{};
  p.idx = idx;
  p.cnt = cnt;
  p.idx_stride = idx_stride;
  p.out = static_cast<uint16_t*>(out->ptr);
  p.out_sb = out->stride[0];
  p.out_sh = out->stride[1];
  p.out_ss = out->stride[2];
  p.lse = lse;
  p.dst_row = dst_row;
  static const bool static_sched = getenv("BLADE_STATIC_SCHED") && atoi(getenv("BLADE_STATIC_SCHED")) != 0;  // A/B knob
  p.sched = workspace && ws_bytes >= kSchedBytes && !static_sched ? static_cast<int*>(workspace) : nullptr;
  p.park = workspace ? reinterpret_cast<uint4*>(static_cast<uint8_t*>(workspace) + kSchedBytes) : nullptr;
  p.B = (int)B;
  p.H = (int)H;
  p.S = (int)S;
  p.Sk = (int)Sk;
  p.nq = nq;
  p.nk = nk;
  p.n_pool = n_pool;
  p.n_pool_tiles = pooled ? (int)ceil_div(n_pool, kBlockN) : 0;
  p.pairs_per_head = (nq + 1) / 2;
  {
    // tail: the last pairs are issued as solo tiles with the KV list split across the two streams, so the last
    // claims are half-length items (BLADE_NO_SPLIT=1 disables, for A/B timing).  With the dynamic queue that is
    // the last G/2 pairs (one solo per CTA); with the static sequence only a leftover round that fills at most half
    // of the CTAs.
    static const bool no_split = getenv("BLADE_NO_SPLIT") && atoi(getenv("BLADE_NO_SPLIT")) != 0;
    const int np = (int)(B * H) * p.pairs_per_head, G = device_sm_count();
    // + the very last claims as HALF tiles (KV sequence of one query tile split across two CTAs, four items per pair):
    // quarter-length items flatten the end of the schedule.  Needs the dynamic queue and the workspace (slots, arrival
    // counters).  BLADE_NO_XSPLIT=1 disables (A/B).
    static const bool no_xsplit = getenv("BLADE_NO_XSPLIT") && atoi(getenv("BLADE_NO_XSPLIT")) != 0;
    // (not with 64x64 quadrant masks: there a half's entries can all be masked for a row, and 0 / 0 would enter the fold)
    const bool can_x = !no_split && !no_xsplit && !multi && !sub64 && p.sched && workspace && ws_bytes >= attn_park_bytes(D);
    static const int env_xs = getenv("BLADE_XSPLIT_PAIRS") ? atoi(getenv("BLADE_XSPLIT_PAIRS")) : -1;   // tuning knobs
    static const int env_solo = getenv("BLADE_SOLO_PAIRS") ? atoi(getenv("BLADE_SOLO_PAIRS")) : -1;
    int split = 0, xsplit = 0;
    plan_tail(np, G, no_split, can_x, multi != nullptr, p.sched != nullptr, attn_max_split_tiles() / 2, env_xs, env_solo, split,
              xsplit);
    p.num_pair_items = np - split - xsplit;
    p.num_solo_pairs = split;
    p.num_split_pairs = xsplit;
    p.num_items = p.num_pair_items + 2 * split + 4 * xsplit;
    if (xsplit) {
      p.split_cnt = static_cast<int*>(workspace) + 16;
      p.split_ws = static_cast<uint8_t*>(workspace) + kSchedBytes + attn_park_region(D);
    }
  }
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.gap = (float)sample_gap;
  p.log_gap_r = pooled ? round_host(logf(round_host((float)sample_gap, bf)), bf) : 0.f;
  p.exact_merge = exact_merge;
  p.sub64 = sub64;
  if (peers && peers->out[0]) {
    BLADE_REQUIRE(peers->n_peers >= 1 && peers->n_peers <= BLADE_MAX_PEERS && peers->my_peer >= 0 &&
                      peers->my_peer < peers->n_peers && peers->rows_per_peer >= 1 && B == 1,
                  BLADE_ERR_ARG, "BladePeers: bad n_peers / my_peer / rows_per_peer (B must be 1)");
    BLADE_REQUIRE(static_cast<int64_t>(peers->n_peers) * peers->rows_per_peer >= S, BLADE_ERR_SHAPE,
                  "peers hold %lld output rows, sequence has %lld", (long long)peers->n_peers * peers->rows_per_peer,
                  (long long)S);
    for (int i = 0; i < peers->n_peers; ++i) {
      BLADE_REQUIRE(peers->out[i] && (reinterpret_cast<uintptr_t>(peers->out[i]) & 15) == 0, BLADE_ERR_ARG,
                    "peer %d: output pointer null or misaligned", i);
      p.out_peer[i] = static_cast<uint16_t*>(peers->out[i]) + static_cast<int64_t>(peers->my_peer) * H * D;
    }
    p.out_peer_rows = peers->rows_per_peer;
    p.out_sh = D;                                       // [rows, H_total, D] token-major
    p.out_ss = static_cast<int64_t>(peers->n_peers) * H * D;
  }
#ifdef BLADE_TRACE
  {
    static long long* tbuf = nullptr;  // debug build only (tools/trace_attn.py); the product never allocates
    if (!tbuf) cudaMalloc(&tbuf, 4 * 8 * 256 * sizeof(long long));
    cudaMemsetAsync(tbuf, 0, 4 * 8 * 256 * sizeof(long long), stream);
    p.trace = tbuf;
    g_trace_buf = tbuf;
  }
#endif

  const int grid = p.num_items < device_sm_count() ? p.num_items : device_sm_count();
  if (p.sched && !sched_prezeroed) BLADE_CUDA_OK(cudaMemsetAsync(p.sched, 0, kSchedBytes, stream));
  StageTimer timer(3, stream);
  if (multi) {
    BLADE_REQUIRE(!pooled && !sub64 && multi->cnt4, BLADE_ERR_ARG, "multi-level attention: no pooled branch / block 64");
    BLADE_REQUIRE((reinterpret_cast<uintptr_t>(multi->cnt4) & 15) == 0, BLADE_ERR_ALIGN, "cnt4 not 16B aligned");
    MultiMaps mm;
    for (int l = 0; l < 3; ++l) {
      const int rows = kBlockN >> (l + 1);
      for (const BladeTensor* t : {multi->k[l], multi->v[l]}) {
        if (int e = check_tensor16(t, "pyramid level")) return e;
        BLADE_REQUIRE(t->shape[0] == B && t->shape[1] == H && t->shape[3] == D && t->shape[2] >= static_cast<int64_t>(nk) * rows &&
                          t->dtype == q->dtype,
                      BLADE_ERR_SHAPE, "pyramid level %d: need [B,H,>=%lld,D]", 2 << l, (long long)nk * rows);
      }
      if (int e = make_tmap(&mm.k[l], multi->k[l], rows)) return e;
      if (int e = make_tmap(&mm.v[l], multi->v[l], rows)) return e;
    }
    p.cnt4 = reinterpret_cast<const int4*>(multi->cnt4);
#define LAUNCH_MULTI(DD, BF)                                                                                         \
  do {                                                                                                               \
    auto kern = asa_multilevel_attn_kernel<DD, BF>;                                                                  \
    BLADE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<DD>::kTotal));  \
    kern<<<grid, kThreads, SmemLayout<DD>::kTotal, stream>>>(tmQ, tmK, tmV, mm, p);                                   \
  } while (0)
    if (D == 128) { if (bf) LAUNCH_MULTI(128, true); else LAUNCH_MULTI(128, false); }
    else          { if (bf) LAUNCH_MULTI(64, true); else LAUNCH_MULTI(64, false); }
#undef LAUNCH_MULTI
    BLADE_CUDA_OK(cudaGetLastError());
    return BLADE_OK;
  }
#define LAUNCH_ATTN(DD, BF)                                                                                     \
  do {                                                                                                          \
    auto kern = pooled ? asa_attn_kernel<DD, BF, true> : asa_attn_kernel<DD, BF, false>;                        \
    BLADE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<DD>::kTotal)); \
    kern<<<grid, kThreads, SmemLayout<DD>::kTotal, stream>>>(tmQ, tmK, tmV, tmKp, tmVp, p);                      \
  } while (0)
  if (D == 128) { if (bf) LAUNCH_ATTN(128, true); else LAUNCH_ATTN(128, false); }
  else          { if (bf) LAUNCH_ATTN(64, true); else LAUNCH_ATTN(64, false); }
#undef LAUNCH_ATTN
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

int launch_probe(const void* a_tile, const void* b_tile, const float* p_in, float* out, int D, int mode,
                 cudaStream_t stream) {
  BLADE_REQUIRE(D == 64 || D == 128, BLADE_ERR_SHAPE, "probe D");
  CUtensorMap tmA, tmB;
  if (int e = make_tmap(&tmA, mode == 0 ? a_tile : b_tile, BLADE_BF16, 1, 1, 128, D, 128 * D, 128 * D, D)) return e;
  if (int e = make_tmap(&tmB, b_tile, BLADE_BF16, 1, 1, 128, D, 128 * D, 128 * D, D)) return e;
  const int smem = 2 * kBlockN * D * 2 + 1024 + 256;
  if (D == 128) {
    BLADE_CUDA_OK(cudaFuncSetAttribute(probe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<128><<<1, 128, smem, stream>>>(tmA, tmB, p_in, out, mode);
  } else {
    BLADE_CUDA_OK(cudaFuncSetAttribute(probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<64><<<1, 128, smem, stream>>>(tmA, tmB, p_in, out, mode);
  }
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

}  // namespace blade

using namespace blade;

extern "C" int blade_block_sparse_attn_fwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                                           const int32_t* idx, const int32_t* cnt, int64_t idx_stride,
                                           BladeTensor* out, float* lse, const int32_t* dst_row, float softmax_scale,
                                           void* workspace, size_t ws_bytes, void* stream) {
  return launch_attn(q, k, v, idx, cnt, idx_stride, nullptr, nullptr, 0, out, lse, dst_row, softmax_scale, 0, workspace,
                     ws_bytes, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

// block_size 64 variants: idx entries = (128-key tile id) | (quadrant mask << 28), built by blade_mask64_to_index
extern "C" int blade_block_sparse_attn64_fwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                                             const int32_t* idx, const int32_t* cnt, int64_t idx_stride,
                                             BladeTensor* out, float* lse, const int32_t* dst_row, float softmax_scale,
                                             void* workspace, size_t ws_bytes, void* stream) {
  g_sub64_next = 1;
  return launch_attn(q, k, v, idx, cnt, idx_stride, nullptr, nullptr, 0, out, lse, dst_row, softmax_scale, 0, workspace,
                     ws_bytes, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}
extern "C" int blade_asa_attn64_fwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v, const int32_t* idx,
                                    const int32_t* cnt, int64_t idx_stride, const BladeTensor* k_pool,
                                    const BladeTensor* v_pool, int32_t sample_gap, BladeTensor* out,
                                    const int32_t* dst_row, float softmax_scale, int32_t exact_merge, void* workspace,
                                    size_t ws_bytes, void* stream) {
  BLADE_REQUIRE(sample_gap > 0 && k_pool && v_pool, BLADE_ERR_ARG, "pooled branch inputs missing");
  g_sub64_next = 1;
  return launch_attn(q, k, v, idx, cnt, idx_stride, k_pool, v_pool, sample_gap, out, nullptr, dst_row, softmax_scale,
                     exact_merge, workspace, ws_bytes, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

extern "C" int blade_asa_attn_fwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v, const int32_t* idx,
                                  const int32_t* cnt, int64_t idx_stride, const BladeTensor* k_pool,
                                  const BladeTensor* v_pool, int32_t sample_gap, BladeTensor* out,
                                  const int32_t* dst_row, float softmax_scale, int32_t exact_merge, void* workspace,
                                  size_t ws_bytes, void* stream) {
  BLADE_REQUIRE(sample_gap > 0 && k_pool && v_pool, BLADE_ERR_ARG, "pooled branch inputs missing");
  return launch_attn(q, k, v, idx, cnt, idx_stride, k_pool, v_pool, sample_gap, out, nullptr, dst_row, softmax_scale,
                     exact_merge, workspace, ws_bytes, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

// Host-side replay of a launch's work-item decode: the SAME classify / pair_id_of / item_counts / make_item code the three
// device roles run, on host counts -- so tests can check, without a GPU, that every query tile's pooled tiles and block
// list are covered exactly once for any (B, H, nq, counts, SM count), half tiles included.
extern "C" int blade_debug_attn_schedule(int64_t B, int64_t H, int64_t nq, int32_t n_pool_tiles, int32_t sm_count,
                                         const int32_t* cnt_host, int32_t dynamic_queue, int32_t half_tiles,
                                         int32_t* items_out, int64_t max_items, int32_t* n_items_out) {
  BLADE_REQUIRE(B >= 1 && H >= 1 && nq >= 1 && sm_count >= 1 && n_pool_tiles >= 0 && cnt_host && items_out && n_items_out,
                BLADE_ERR_ARG, "bad argument");
  AttnParams p{};
  p.B = static_cast<int>(B);
  p.H = static_cast<int>(H);
  p.nq = static_cast<int>(nq);
  p.n_pool_tiles = n_pool_tiles;
  p.pairs_per_head = static_cast<int>((nq + 1) / 2);
  p.cnt = cnt_host;
  const int np = static_cast<int>(B * H) * p.pairs_per_head, G = sm_count;
  const int xmax = (G < kMaxSplitTiles ? G : kMaxSplitTiles) / 2;
  int split = 0, xsplit = 0;
  plan_tail(np, G, false, dynamic_queue != 0 && half_tiles != 0, false, dynamic_queue != 0, xmax, -1, -1, split, xsplit);
  p.num_pair_items = np - split - xsplit;
  p.num_solo_pairs = split;
  p.num_split_pairs = xsplit;
  p.num_items = p.num_pair_items + 2 * split + 4 * xsplit;
  *n_items_out = p.num_items;
  BLADE_REQUIRE(p.num_items <= max_items, BLADE_ERR_WORKSPACE, "%d items, room for %lld", p.num_items, (long long)max_items);
  for (int item = 0; item < p.num_items; ++item) {
    int c0, c1;
    item_counts(p, item, c0, c1);
    const Item it = make_item(p, item, c0, c1);
    int32_t* o = items_out + static_cast<int64_t>(item) * 12;
    o[0] = item;
    o[1] = it.bh;
    for (int t = 0; t < 2; ++t) {
      o[2 + 4 * t] = it.qb[t];
      o[3 + 4 * t] = it.pt[t];
      o[4 + 4 * t] = it.off[t];
      o[5 + 4 * t] = it.ns[t];
    }
    o[10] = it.merge ? 1 : 0;
    o[11] = it.split ? 1 + 2 * it.slot + it.half : 0;
  }
  return BLADE_OK;
}

#ifdef BLADE_TRACE
extern "C" int blade_debug_trace(long long* host_out) {
  if (!g_trace_buf) return 1;
  return cudaMemcpy(host_out, g_trace_buf, 4 * 8 * 256 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess;
}
#endif

extern "C" int blade_probe_qk(const void* q_tile, const void* k_tile, float* s_out, int32_t D, void* stream) {
  return launch_probe(q_tile, k_tile, nullptr, s_out, D, 0, static_cast<cudaStream_t>(stream));
}
extern "C" int blade_probe_pv(const float* p_tile, const void* v_tile, float* o_out, int32_t D, void* stream) {
  return launch_probe(nullptr, v_tile, p_tile, o_out, D, 1, static_cast<cudaStream_t>(stream));
}
