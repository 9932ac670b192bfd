// common.cuh -- host-side error plumbing and small shared helpers for the blade_asa C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/blade_asa.h"

namespace blade {

char* last_error_buf();  // thread-local, 512 bytes (defined in capi.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define BLADE_CUDA_OK(expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::blade::fail(BLADE_ERR_LAUNCH, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                 \
  } while (0)

#define BLADE_REQUIRE(cond, code, ...) \
  do {                                 \
    if (!(cond)) return ::blade::fail(code, __VA_ARGS__); \
  } while (0)

// optional profiling events (blade_profile_events): stage 0 prep, 1 scores, 2 select, 3 attention, 4 gap pooling
struct StageEvents {
  cudaEvent_t start, stop;
};
StageEvents* stage_events();  // array of 5 (defined in capi.cu)
struct StageTimer {
  int stage;
  cudaStream_t stream;
  StageTimer(int s, cudaStream_t st) : stage(s), stream(st) {
    if (stage >= 0 && stage_events()[stage].start) cudaEventRecord(stage_events()[stage].start, stream);
  }
  ~StageTimer() {
    if (stage >= 0 && stage_events()[stage].stop) cudaEventRecord(stage_events()[stage].stop, stream);
  }
};

// fused q/k RMSNorm (BladeQkNorm): rstd = device fp32 scratch [2][B*S] that prep_impl fills itself
struct PrepNorm {
  int kind;
  float eps;
  const void* q_weight;
  const void* k_weight;
  float* rstd;            // scratch the statistic kernel fills ...
  const float* rstd_ext;  // ... unless the caller brings the statistic (indexed by token)
  const void* q_bias;     // kind 3 (per-head LayerNorm) only
  const void* k_bias;
};
int prep_impl(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v, const int32_t* src_row, void* q_r,
              void* k_r, void* v_r, float* q_mean, float* k_mean, void* k_pool, void* v_pool, int32_t block_size,
              int32_t sample_gap, const float* rope_cos_sin, int32_t rope_first_row, int parts, cudaStream_t stream,
              const PrepNorm* norm = nullptr, const int32_t* tok_row = nullptr, const BladePeers* peers = nullptr,
              int tmask = 7 /* bit t: tensor t of (q, k, v) */, int stage = -2 /* profile-event stage override */);
int score_select_impl(const float* q_mean, const float* k_mean, float* scores_opt, int64_t B, int64_t H, int64_t nb,
                      int64_t D, const BladeAsaConfig* cfg, int32_t* idx, int32_t* cnt, uint8_t* mask_opt,
                      cudaStream_t stream, bool pdl);
int rms_stat_impl(const BladeTensor* q, const BladeTensor* k, float eps, float* out, cudaStream_t stream,
                  float* const* peer_out = nullptr, int n_peers = 0, int64_t out_rows = 0, int64_t out_row0 = 0);

// ---- attention launch (attn_kernel.cu)
// multi-level launch: the pooled K/V pyramid (levels 2, 4, 8: [B,H,nk*128/L,D] each) and the per-level entry counts
struct MultiLevelArgs {
  const BladeTensor* k[3];
  const BladeTensor* v[3];
  const int32_t* cnt4;  // device int32 [B,H,nq,4]: list entries of level 1, 2, 4, 8 (list sorted by level, then block id)
};
int device_sm_count();  // SMs of the current device (cached per device)
size_t attn_park_bytes(int64_t D);
size_t attn_sched_bytes();  // head of the attention workspace that must be zero at launch (item + arrival counters)
void attn_sched_prezeroed();
void attn_next_sub64();
int launch_attn(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v, const int32_t* idx,
                const int32_t* cnt, int64_t idx_stride, const BladeTensor* k_pool, const BladeTensor* v_pool,
                int32_t sample_gap, BladeTensor* out, float* lse, const int32_t* dst_row, float softmax_scale,
                int exact_merge, void* workspace, size_t ws_bytes, cudaStream_t stream,
                const BladePeers* peers = nullptr, const MultiLevelArgs* multi = nullptr);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// validates a [B,H,S,D] 16-bit tensor for the vectorised / TMA paths
inline int check_tensor16(const BladeTensor* t, const char* name) {
  BLADE_REQUIRE(t && t->ptr, BLADE_ERR_ARG, "%s: null tensor", name);
  BLADE_REQUIRE(t->dtype == BLADE_BF16 || t->dtype == BLADE_F16, BLADE_ERR_DTYPE,
                "%s: dtype %d unsupported (bf16/f16 only)", name, t->dtype);
  BLADE_REQUIRE(t->stride[3] == 1, BLADE_ERR_ALIGN, "%s: last dim must be contiguous", name);
  BLADE_REQUIRE((reinterpret_cast<uintptr_t>(t->ptr) & 15) == 0, BLADE_ERR_ALIGN, "%s: pointer not 16B aligned", name);
  for (int i = 0; i < 3; ++i)
    BLADE_REQUIRE((t->stride[i] * 2) % 16 == 0, BLADE_ERR_ALIGN, "%s: stride[%d]=%lld not 16B aligned", name, i,
                  (long long)t->stride[i]);
  BLADE_REQUIRE(t->shape[3] == 64 || t->shape[3] == 128, BLADE_ERR_SHAPE, "%s: head dim %lld not in {64,128}", name,
                (long long)t->shape[3]);
  return BLADE_OK;
}

}  // namespace blade
