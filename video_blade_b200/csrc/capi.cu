// capi.cu -- C ABI glue: error state, device check, Gilbert tables (host), workspace carving and the
// whole-layer entry point blade_asa_forward (AdaptiveBlockSparseAttnTrain.forward, W:383-408 / C:405-427).
#include <math.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include <stdlib.h>

#include "common.cuh"

namespace blade {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

StageEvents* stage_events() {
  static thread_local StageEvents ev[5] = {};  // measurement hook: per calling thread, like the launches it brackets
  return ev;
}

// ---------------------------------------------------------------------------------------------
// Generalised Hilbert curve (the algorithm of gilbert3d.py:6-167), iterative, integer 3-vectors.
// ---------------------------------------------------------------------------------------------
struct V3 {
  int x, y, z;
};
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline int sgn(int v) { return (v > 0) - (v < 0); }
static inline V3 sgn(V3 a) { return {sgn(a.x), sgn(a.y), sgn(a.z)}; }
static inline int fdiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }  // floor(v/2)
static inline V3 half(V3 a) { return {fdiv2(a.x), fdiv2(a.y), fdiv2(a.z)}; }
static inline int mag(V3 a) { return abs(a.x + a.y + a.z); }

struct Box {
  V3 p, a, b, c;
};

static void gilbert_fill(int W, int H, int D, int64_t* curve2raster) {
  std::vector<Box> stack;
  const V3 o{0, 0, 0}, ex{W, 0, 0}, ey{0, H, 0}, ez{0, 0, D};
  if (W >= H && W >= D) stack.push_back({o, ex, ey, ez});
  else if (H >= W && H >= D) stack.push_back({o, ey, ex, ez});
  else stack.push_back({o, ez, ex, ey});
  int64_t n = 0;
  auto emit_run = [&](V3 p, V3 d, int len) {
    for (int i = 0; i < len; ++i) {
      curve2raster[n++] = p.x + static_cast<int64_t>(W) * (p.y + static_cast<int64_t>(H) * p.z);
      p = p + d;
    }
  };
  while (!stack.empty()) {
    const Box bx = stack.back();
    stack.pop_back();
    const V3 p = bx.p, a = bx.a, b = bx.b, c = bx.c;
    const int w = mag(a), h = mag(b), d = mag(c);
    const V3 da = sgn(a), db = sgn(b), dc = sgn(c);
    if (h == 1 && d == 1) { emit_run(p, da, w); continue; }
    if (w == 1 && d == 1) { emit_run(p, db, h); continue; }
    if (w == 1 && h == 1) { emit_run(p, dc, d); continue; }
    V3 a2 = half(a), b2 = half(b), c2 = half(c);
    const int w2 = mag(a2), h2 = mag(b2), d2 = mag(c2);
    if ((w2 % 2) && w > 2) a2 = a2 + da;
    if ((h2 % 2) && h > 2) b2 = b2 + db;
    if ((d2 % 2) && d > 2) c2 = c2 + dc;
    Box kids[5];
    int nk = 0;
    if (2 * w > 3 * h && 2 * w > 3 * d) {
      kids[nk++] = {p, a2, b, c};
      kids[nk++] = {p + a2, a - a2, b, c};
    } else if (3 * h > 4 * d) {
      kids[nk++] = {p, b2, c, a2};
      kids[nk++] = {p + b2, a, b - b2, c};
      kids[nk++] = {p + (a - da) + (b2 - db), -b2, c, -(a - a2)};
    } else if (3 * d > 4 * h) {
      kids[nk++] = {p, c2, a2, b};
      kids[nk++] = {p + c2, a, b, c - c2};
      kids[nk++] = {p + (a - da) + (c2 - dc), -c2, -(a - a2), b};
    } else {
      kids[nk++] = {p, b2, c2, a2};
      kids[nk++] = {p + b2, c, a2, b - b2};
      kids[nk++] = {p + (b2 - db) + (c - dc), a, -b2, -(c - c2)};
      kids[nk++] = {p + (a - da) + b2 + (c - dc), -c, -(a - a2), b - b2};
      kids[nk++] = {p + (a - da) + (b2 - db), -b2, c2, -(a - a2)};
    }
    for (int i = nk - 1; i >= 0; --i) stack.push_back(kids[i]);
  }
}

// side stream + events for the fork/join inside blade_asa_forward: one set PER CALLER STREAM (created on first use, kept
// for the life of the process), so that layer calls in flight on different streams -- the two CFG branches of a
// sampler step, or two host threads -- never share a side stream or an event.  Calls on the SAME stream are ordered by
// the stream itself.  The side stream has the HIGHEST priority: it carries short kernels whose CTAs the block
// scheduler then places ahead of the pending CTAs of a grid-filling, DRAM-bound kernel on the caller's stream.
struct ForkState {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static ForkState* fork_state(cudaStream_t caller) {
  static std::mutex mu;
  static std::unordered_map<uint64_t, ForkState> table;
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t key = (static_cast<uint64_t>(dev) << 56) ^ reinterpret_cast<uint64_t>(caller);
  std::lock_guard<std::mutex> lock(mu);
  ForkState& f = table[key];
  if (!f.side) {
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    cudaStreamCreateWithPriority(&f.side, cudaStreamNonBlocking, greatest);
    cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming);
  }
  return &f;   // unordered_map never moves its nodes: the pointer stays valid
}

// workspace carving for blade_asa_forward
struct ForwardWs {
  size_t q_r, k_r, v_r, q_mean, k_mean, k_pool, v_pool, scores, idx, cnt, park, rstd, q_s, k_s, mask64, idx128, cnt128,
      total;
};
static ForwardWs carve(int64_t B, int64_t H, int64_t S, int64_t D, const BladeAsaConfig* cfg) {
  ForwardWs w{};
  const int64_t blk = cfg->block_size > 0 ? cfg->block_size : 128;
  const int64_t nb = ceil_div(S, blk);
  const int64_t np = cfg->sample_gap > 0 ? ceil_div(S, cfg->sample_gap) : 0;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  w.q_r = take(B * H * S * D * 2);
  w.k_r = take(B * H * S * D * 2);
  w.v_r = take(B * H * S * D * 2);
  w.q_mean = take(B * H * nb * D * 4);
  w.k_mean = take(B * H * nb * D * 4);
  w.k_pool = take(B * H * (np > 0 ? np : 1) * D * 2);
  w.v_pool = take(B * H * (np > 0 ? np : 1) * D * 2);
  w.scores = take(B * H * nb * nb * 4);
  w.idx = take(B * H * nb * nb * 4);
  w.cnt = take(B * H * nb * 4);
  w.park = take(attn_park_bytes(D));
  w.rstd = take(2 * B * S * 4);  // q/k RMSNorm statistic (BladeQkNorm)
  if (cfg->estimator == 1) {     // sampled-max estimator: the num_keep sampled rows of every block (W:37-60)
    w.q_s = take(B * H * nb * 32 * D * 2);
    w.k_s = take(B * H * nb * 32 * D * 2);
  }
  if (blk == 64) {               // 64-granular selection feeding the 128x128 tensor-core tiles
    const int64_t n128 = ceil_div(nb, 2);
    w.mask64 = take(B * H * nb * nb);
    w.idx128 = take(B * H * n128 * n128 * 4);
    w.cnt128 = take(B * H * n128 * 4);
  }
  w.total = off;
  return w;
}

}  // namespace blade

using namespace blade;

static_assert(sizeof(BladeAsaConfig) == 104 && sizeof(BladeTensor) == 80 && sizeof(BladePeers) == 16 + 4 * 8 * BLADE_MAX_PEERS,
              "C ABI struct layout changed: bump BLADE_ABI_VERSION and the ctypes mirrors in video_blade_b200/_lib.py");

extern "C" int blade_abi_version(void) { return BLADE_ABI_VERSION; }
extern "C" const char* blade_last_error(void) { return last_error_buf(); }

extern "C" int blade_profile_events(int32_t stage, void* start_event, void* stop_event) {
  BLADE_REQUIRE(stage >= 0 && stage < 5, BLADE_ERR_ARG, "stage %d out of range", stage);
  stage_events()[stage].start = static_cast<cudaEvent_t>(start_event);
  stage_events()[stage].stop = static_cast<cudaEvent_t>(stop_event);
  return BLADE_OK;
}

extern "C" int blade_device_check(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(BLADE_ERR_NO_DEVICE, "no CUDA device (%s); blade_asa has no CPU path", cudaGetErrorString(e));
  }
  int dev = 0, major = 0;
  BLADE_CUDA_OK(cudaGetDevice(&dev));
  BLADE_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  BLADE_REQUIRE(major == 10, BLADE_ERR_NO_DEVICE, "device compute capability %d.x is not sm_100", major);
  return BLADE_OK;
}

extern "C" int blade_gilbert_tables(int32_t width, int32_t height, int32_t depth, int64_t* curve2raster,
                                    int64_t* raster2curve) {
  BLADE_REQUIRE(width >= 1 && height >= 1 && depth >= 1, BLADE_ERR_ARG, "grid dims must be >= 1");
  BLADE_REQUIRE(curve2raster, BLADE_ERR_ARG, "curve2raster null");
  const int64_t n = static_cast<int64_t>(width) * height * depth;
  gilbert_fill(width, height, depth, curve2raster);
  if (raster2curve)
    for (int64_t c = 0; c < n; ++c) raster2curve[curve2raster[c]] = c;
  return BLADE_OK;
}

extern "C" size_t blade_attn_workspace_bytes(int64_t D) { return attn_park_bytes(D); }

extern "C" size_t blade_asa_workspace_bytes(int64_t B, int64_t H, int64_t S, int64_t D, const BladeAsaConfig* cfg) {
  if (!cfg) return 0;
  return carve(B, H, S, D, cfg).total;
}

extern "C" int blade_asa_forward(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                                 const int32_t* src_row, const int32_t* dst_row, const BladeAsaConfig* cfg,
                                 const float* scores_in, BladeTensor* out, float* scores_out, uint8_t* mask_out,
                                 int32_t* idx_out, int32_t* cnt_out, void* workspace, size_t ws_bytes,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BLADE_REQUIRE(cfg, BLADE_ERR_ARG, "cfg null");
  if (int e = check_tensor16(q, "q")) return e;
  BLADE_REQUIRE(cfg->block_size == 128 || cfg->block_size == 64, BLADE_ERR_ARG,
                "blade_asa_forward: block size %d not in {64, 128} (W:325)", cfg->block_size);
  BLADE_REQUIRE(cfg->estimator == 0 || cfg->estimator == 1, BLADE_ERR_ARG, "estimator %d not in {0 = mean-pool, 1 = sampled-max}",
                cfg->estimator);
  const bool sampled = cfg->estimator == 1 && !scores_in;
  BLADE_REQUIRE(!sampled || (cfg->sample_q_off && cfg->sample_k_off), BLADE_ERR_ARG,
                "estimator 1 (sampled-max, W:62-87) needs cfg->sample_q_off / sample_k_off (device int32 [B,H,32])");
  BLADE_REQUIRE(!sampled || cfg->num_keep == 32, BLADE_ERR_ARG, "the sampled estimator is built for num_keep = 32 (W:62)");
  const bool blk64 = cfg->block_size == 64;
  const int64_t B = q->shape[0], H = q->shape[1], S = q->shape[2], D = q->shape[3];
  const ForwardWs w = carve(B, H, S, D, cfg);
  BLADE_REQUIRE(workspace && ws_bytes >= w.total, BLADE_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu",
                w.total, ws_bytes);
  BLADE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, BLADE_ERR_ALIGN, "workspace must be 1 KiB aligned");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int64_t nb = ceil_div(S, cfg->block_size);
  const int64_t np = cfg->sample_gap > 0 ? ceil_div(S, cfg->sample_gap) : 0;
  const bool norm_on = cfg->qk_norm != nullptr && cfg->qk_norm->kind != 0;
  const bool pull = cfg->peers != nullptr && cfg->peers->q[0] != nullptr;  // q/k/v rows live in the peers' memory
  const bool rearr = src_row != nullptr || cfg->rope_cos_sin != nullptr || norm_on || pull;  // these need the output copies
  const bool v_copy = src_row != nullptr || pull;  // rope / norm rewrite q and k only: v is then read in place
  PrepNorm pn{};
  if (norm_on) {
    pn.kind = cfg->qk_norm->kind;
    pn.eps = cfg->qk_norm->eps;
    pn.q_weight = cfg->qk_norm->q_weight;
    pn.k_weight = cfg->qk_norm->k_weight;
    pn.rstd = reinterpret_cast<float*>(ws + w.rstd);
    pn.rstd_ext = cfg->qk_norm->rstd;
    pn.q_bias = cfg->qk_norm->q_bias;
    pn.k_bias = cfg->qk_norm->k_bias;
  }
  float* q_mean = reinterpret_cast<float*>(ws + w.q_mean);
  float* k_mean = reinterpret_cast<float*>(ws + w.k_mean);
  float* scores = scores_out ? scores_out : reinterpret_cast<float*>(ws + w.scores);
  int32_t* idx = idx_out ? idx_out : reinterpret_cast<int32_t*>(ws + w.idx);
  int32_t* cnt = cnt_out ? cnt_out : reinterpret_cast<int32_t*>(ws + w.cnt);

  // the attention kernel's item counter (head of its workspace) is zeroed here, ahead of the mask kernels, so the
  // memset is not on the path between the selection and the attention launch
  BLADE_CUDA_OK(cudaMemsetAsync(ws + w.park, 0, attn_sched_bytes(), stream));
  const bool need_means = scores_in == nullptr && !sampled;
  // gather / rotate / block means on the caller's stream ...
  // q and k first (the score + selection chain only needs their block means); v's copy runs on the side stream in front
  // of the pooling, next to the latency-bound score / selection kernel.  BLADE_SPLIT_V=0 restores one launch (A/B).
  static const bool split_v = !(getenv("BLADE_SPLIT_V") && atoi(getenv("BLADE_SPLIT_V")) == 0);
  static const int fork_mode0 = getenv("BLADE_FORK_MODE") ? atoi(getenv("BLADE_FORK_MODE")) : 0;
  const bool v_later = split_v && v_copy && np > 0 && fork_mode0 == 0;
  if (int e = prep_impl(q, k, v, src_row, rearr ? ws + w.q_r : nullptr, rearr ? ws + w.k_r : nullptr,
                        v_copy ? ws + w.v_r : nullptr, need_means ? q_mean : nullptr, need_means ? k_mean : nullptr,
                        nullptr, nullptr, cfg->block_size, cfg->sample_gap, cfg->rope_cos_sin, cfg->rope_first_row, 1,
                        stream, norm_on ? &pn : nullptr, cfg->token_row, cfg->peers, v_later ? 3 : 7))
    return e;
  // ... then the bandwidth-bound gap pooling and the latency-bound score + selection kernels run concurrently on
  // two streams, joined right before the attention launch (events only, no host sync).  BLADE_FORK_MODE (A/B knob):
  // 0 = pooling on the side stream, 1 = score/selection on the (high-priority) side stream, 2 = no fork.
  const int fork_mode = fork_mode0;
  ForkState* fk = (np && fork_mode != 2) ? fork_state(stream) : nullptr;
  cudaStream_t mstream = stream, pstream = stream;  // streams of the mask kernels / the pooling kernel
  if (fk) {
    if (fork_mode == 1) mstream = fk->side; else pstream = fk->side;
    BLADE_CUDA_OK(cudaEventRecord(fk->fork, stream));
    BLADE_CUDA_OK(cudaStreamWaitEvent(fk->side, fk->fork, 0));
  }
  auto run_pool = [&]() -> int {
    StageTimer timer(4, pstream);   // stage 4 = everything on the side stream: (v's copy +) gap pooling
    if (v_later)
      if (int e = prep_impl(q, k, v, src_row, ws + w.q_r, ws + w.k_r, ws + w.v_r, nullptr, nullptr, nullptr, nullptr,
                            cfg->block_size, cfg->sample_gap, nullptr, 0, 1, pstream, nullptr, cfg->token_row, cfg->peers,
                            4, -1))
        return e;
    return prep_impl(q, k, v, src_row, rearr ? ws + w.q_r : nullptr, rearr ? ws + w.k_r : nullptr,
                     v_copy ? ws + w.v_r : nullptr, nullptr, nullptr, ws + w.k_pool, ws + w.v_pool, cfg->block_size,
                     cfg->sample_gap, nullptr, 0, 2, pstream, nullptr, nullptr, nullptr, 7, -1);
  };
  if (np && fork_mode != 1)
    if (int e = run_pool()) return e;
  BladeTensor qr = *q, kr = *k, vr = *v;
  if (rearr) {
    const int64_t cs[4] = {H * S * D, S * D, D, 1};
    qr.ptr = ws + w.q_r;
    kr.ptr = ws + w.k_r;
    for (int i = 0; i < 4; ++i) qr.stride[i] = kr.stride[i] = cs[i];
    if (v_copy) {
      vr.ptr = ws + w.v_r;
      for (int i = 0; i < 4; ++i) vr.stride[i] = cs[i];
    }
  }
  const float* sc = scores_in;
  bool fused_select = false;
  if (sampled) {
    // the reference's estimator (efficient_attn_with_pooling, W:62-87 -> P:201-253) on the gathered / rotated /
    // normalised q and k: sample num_keep rows per block, then the tcgen05 score kernel
    if (int e = blade_asa_sample_tokens(&qr, &kr, cfg->sample_q_off, cfg->sample_k_off, ws + w.q_s, ws + w.k_s,
                                        cfg->block_size, mstream))
      return e;
    if (int e = blade_asa_scores_sampled(ws + w.q_s, ws + w.k_s, scores, B, H, nb, D, q->dtype, mstream)) return e;
    sc = scores;
  } else if (!sc) {
    // mean-pool estimator: block-score GEMM + softmax + selection in ONE launch (the fp32 score map only goes to memory
    // when the caller asked for it); shapes the fused kernel does not cover take the two-kernel path below
    static const bool no_fuse = getenv("BLADE_NO_FUSED_SELECT") && atoi(getenv("BLADE_NO_FUSED_SELECT")) != 0;  // A/B knob
    int rc = -1;
    if (!no_fuse)
      rc = score_select_impl(q_mean, k_mean, scores_out, B, H, nb, D, cfg, blk64 ? reinterpret_cast<int32_t*>(ws + w.idx) : idx,
                             cnt, blk64 ? (mask_out ? mask_out : ws + w.mask64) : mask_out, mstream, false);
    if (rc > 0) return rc;
    if (rc == 0) {
      fused_select = true;
    } else {
      if (int e = blade_asa_scores_meanpool(q_mean, k_mean, scores, B, H, nb, D, mstream)) return e;
      sc = scores;
    }
  } else if (scores_out) {
    BLADE_CUDA_OK(cudaMemcpyAsync(scores_out, scores_in, B * H * nb * nb * 4, cudaMemcpyDeviceToDevice, mstream));
  }
  const int32_t* attn_idx = idx;
  const int32_t* attn_cnt = cnt;
  int64_t attn_stride = nb;
  if (blk64) {
    // block_size 64 (BASELINE config 1): select on the 64-granular score map, then fold the bool mask into
    // quadrant-flagged lists over the 128x128 tensor-core tiles.  cnt_out = the 64-granular counts [B,H,nb];
    // idx_out (if given) = the 128-tile lists [B,H,ceil(nb/2),ceil(nb/2)].
    uint8_t* m64 = mask_out ? mask_out : ws + w.mask64;
    int32_t* i64 = reinterpret_cast<int32_t*>(ws + w.idx);
    int32_t* i128 = idx_out ? idx_out : reinterpret_cast<int32_t*>(ws + w.idx128);
    int32_t* c128 = reinterpret_cast<int32_t*>(ws + w.cnt128);
    if (!fused_select)
      if (int e = blade_asa_select(sc, B, H, nb, nb, cfg, nullptr, nullptr, i64, cnt, m64, nullptr, mstream)) return e;
    if (int e = blade_mask64_to_index(m64, B, H, nb, nb, i128, c128, mstream)) return e;
    attn_idx = i128;
    attn_cnt = c128;
    attn_stride = ceil_div(nb, 2);
  } else if (!fused_select) {
    if (int e = blade_asa_select(sc, B, H, nb, nb, cfg, nullptr, nullptr, idx, cnt, mask_out, nullptr, mstream)) return e;
  }
  if (np && fork_mode == 1)
    if (int e = run_pool()) return e;
  if (fk) {
    BLADE_CUDA_OK(cudaEventRecord(fk->join, fk->side));
    BLADE_CUDA_OK(cudaStreamWaitEvent(stream, fk->join, 0));
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(D));
  if (np) {
    BladeTensor kp{}, vp{};
    kp.ptr = ws + w.k_pool;
    vp.ptr = ws + w.v_pool;
    const int64_t ps[4] = {H * np * D, np * D, D, 1};
    const int64_t sh[4] = {B, H, np, D};
    for (int i = 0; i < 4; ++i) {
      kp.shape[i] = vp.shape[i] = sh[i];
      kp.stride[i] = vp.stride[i] = ps[i];
    }
    kp.dtype = vp.dtype = q->dtype;
    attn_sched_prezeroed();
    if (blk64) attn_next_sub64();
    return launch_attn(&qr, &kr, &vr, attn_idx, attn_cnt, attn_stride, &kp, &vp, cfg->sample_gap, out, nullptr, dst_row,
                       scale, cfg->exact_merge, ws + w.park, attn_park_bytes(D), stream, cfg->peers);
  }
  attn_sched_prezeroed();
  if (blk64) attn_next_sub64();
  return launch_attn(&qr, &kr, &vr, attn_idx, attn_cnt, attn_stride, nullptr, nullptr, 0, out, nullptr, dst_row, scale, 0,
                     ws + w.park, attn_park_bytes(D), stream, cfg->peers);
}
