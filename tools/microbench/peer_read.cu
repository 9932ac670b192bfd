// Microbenchmark (diagnostic only): how fast can one GPU PULL 256-byte rows out of a peer's memory over NVLink?
// Mirrors the Ulysses peer-plane gather (prep_block_kernel<..., PEER>): rows of 256 B at a 3 072 B stride (12 heads x
// 128 channels, token-major), 32 760 x 3/4 rows x 3 heads x 2 tensors ~ 37.7 MB, written to a contiguous local buffer.
//   mode 0: ld.global.nc.L1::no_allocate.v4 (what the kernel does), 8 loads in flight per thread
//   mode 1: plain ld.global.v4, 8 in flight
//   mode 2: like 0 with 16 in flight
//   mode 3: cp.async.bulk (TMA engine) of each 256-B row into shared memory, 128 rows per CTA in flight, then st.global
//   mode 4: contiguous 37.7 MB copy with ld.v4 (upper bound of the link for SM loads)
// usage: peer_read [src_dev=1]     (runs on device 0)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint4 ld_nc(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_plain(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// rows: n_rows rows of 256 B; src row r lives at src + perm(r) * stride_bytes
template <int MODE, int INFLIGHT>
__global__ void __launch_bounds__(256) pull_ld(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n_rows,
                                               int64_t stride_bytes, int rows_total_src) {
  const int lane16 = threadIdx.x & 15, rgrp = threadIdx.x >> 4;        // 16 lanes per row, 16 rows per pass
  const int row0 = blockIdx.x * 128;                                    // 128 rows per CTA = 8 passes
  for (int p0 = 0; p0 < 8; p0 += INFLIGHT) {
    uint4 v[INFLIGHT];
#pragma unroll
    for (int u = 0; u < INFLIGHT; ++u) {
      const int r = row0 + (p0 + u) * 16 + rgrp;
      if (r < n_rows) {
        const int64_t sr = (static_cast<int64_t>(r) * 2654435761u) % rows_total_src;   // scattered source row
        const uint4* p = reinterpret_cast<const uint4*>(src + sr * stride_bytes) + lane16;
        v[u] = MODE == 1 ? ld_plain(p) : ld_nc(p);
      }
    }
#pragma unroll
    for (int u = 0; u < INFLIGHT; ++u) {
      const int r = row0 + (p0 + u) * 16 + rgrp;
      if (r < n_rows) reinterpret_cast<uint4*>(dst + static_cast<int64_t>(r) * 256)[lane16] = v[u];
    }
  }
}

__global__ void __launch_bounds__(128) pull_tma(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n_rows,
                                                int64_t stride_bytes, int rows_total_src) {
  __shared__ __align__(128) uint8_t buf[128 * 256];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_a = static_cast<uint32_t>(__cvta_generic_to_shared(&bar));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int row0 = blockIdx.x * 128;
  const int nr = min(128, n_rows - row0);
  if (threadIdx.x == 0)
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar_a), "r"(nr * 256) : "memory");
  __syncthreads();
  const int r = row0 + threadIdx.x;
  if (threadIdx.x < nr) {
    const int64_t sr = (static_cast<int64_t>(r) * 2654435761u) % rows_total_src;
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(buf + threadIdx.x * 256));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];" ::"r"(d),
                 "l"(src + sr * stride_bytes), "r"(bar_a)
                 : "memory");
  }
  // wait (parity 0)
  asm volatile(
      "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar_a)
      : "memory");
  for (int e = threadIdx.x; e < nr * 16; e += 128)
    reinterpret_cast<uint4*>(dst + static_cast<int64_t>(row0) * 256)[e] = reinterpret_cast<const uint4*>(buf)[e];
}

__global__ void __launch_bounds__(256) copy_contig(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n16) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * 256 * 8 + threadIdx.x;
  uint4 v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (i + u * 256 < n16) v[u] = ld_nc(src + i + u * 256);
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (i + u * 256 < n16) dst[i + u * 256] = v[u];
}

int main(int argc, char** argv) {
  const int src_dev = argc > 1 ? atoi(argv[1]) : 1;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  const int rows_src = 8190, n_rows = 32760 * 3 / 4 * 3 * 2;             // one peer shard; rows pulled (q and k, 3 heads)
  const int64_t stride = 3072, src_bytes = static_cast<int64_t>(rows_src) * stride * 3;
  uint8_t *src = nullptr, *dst = nullptr;
  CK(cudaSetDevice(0));
  if (src_dev != 0) {
    if (src_dev >= ndev) { printf("only %d device(s)\n", ndev); return 0; }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, src_dev));
    if (!can) { printf("no peer access 0 -> %d\n", src_dev); return 0; }
    CK(cudaDeviceEnablePeerAccess(src_dev, 0));
    CK(cudaSetDevice(src_dev));
  }
  CK(cudaMalloc(&src, src_bytes));
  CK(cudaMemset(src, 1, src_bytes));
  CK(cudaDeviceSynchronize());
  CK(cudaSetDevice(0));
  CK(cudaMalloc(&dst, static_cast<int64_t>(n_rows) * 256));
  const int grid = (n_rows + 127) / 128;
  const double mb = n_rows * 256.0 / 1e6;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int rows_total = rows_src * 3;   // rows of `stride` bytes in the source allocation
  for (int mode = 0; mode < 5; ++mode) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      CK(cudaEventRecord(e0));
      if (mode == 0) pull_ld<0, 8><<<grid, 256>>>(src, dst, n_rows, stride, rows_total);
      if (mode == 1) pull_ld<1, 8><<<grid, 256>>>(src, dst, n_rows, stride, rows_total);
      if (mode == 2) pull_ld<0, 4><<<grid, 256>>>(src, dst, n_rows, stride, rows_total);
      if (mode == 3) pull_tma<<<grid, 128>>>(src, dst, n_rows, stride, rows_total);
      if (mode == 4) copy_contig<<<(n_rows * 16 + 2047) / 2048, 256>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), static_cast<int64_t>(n_rows) * 16);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    const char* names[5] = {"ld.nc.no_allocate x8", "ld plain x8", "ld.nc x4 in flight", "cp.async.bulk 256 B rows", "contiguous ld.nc x8"};
    printf("src dev %d  %-26s: %.1f us  %.0f GB/s  (%s)\n", src_dev, names[mode], best * 1e3, mb / best, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
