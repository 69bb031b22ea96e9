// Shared helpers for libstargcn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/stargcn_b200.h"

namespace sg {

// ---- thread-local error string + launch counter (the only state the library keeps) ----
char *err_buf();
void count_launch(int n = 1);
int fail(int code, const char *fmt, ...);

#define SG_REQUIRE(cond, ...)                                      \
  do {                                                             \
    if (!(cond)) return ::sg::fail(SG_ERR_INVALID, __VA_ARGS__);   \
  } while (0)

#define SG_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return ::sg::fail(SG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                  \
  } while (0)

// Checks the launch that was just enqueued (no synchronisation).
#define SG_LAUNCHED(name)                                                                      \
  do {                                                                                         \
    ::sg::count_launch();                                                                      \
    cudaError_t e__ = cudaPeekAtLastError();                                                   \
    if (e__ != cudaSuccess)                                                                    \
      return ::sg::fail(SG_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

inline bool valid_req(int req) { return req == SG_REQ_NULL || req == SG_REQ_WRITE || req == SG_REQ_ADD; }

// SM count of the CURRENT device (cached per device: a process may drive several GPUs)
inline int num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;  // B200
  int n = cache[dev];
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return n;
}

// Development options (sg_dev_option): process-wide integers a tuning tool sets EXPLICITLY through the C ABI —
// the product path never sets them and nothing reads the environment.  Every option's 0 is the shipped behaviour.
enum DevOption {
  SG_DEV_GATHER_VARIANT = 0,   // 1: bulk-copy (TMA) staged index/weight segments, warp per work item (gather.cu)
  SG_DEV_GATHER_GRID = 1,      // blocks per SM of the grid-stride gather launches (0 = 32)
  SG_DEV_GEMM_ARRIVE = 2,      // 1: cluster-scope RELEASE arrival when a TMEM buffer is handed back (0 = relaxed)
  SG_DEV_GEMM_CHAIN = 3,       // k-blocks per TMEM accumulation chain (0 = kChainKBlocks)
  SG_DEV_GATHER_THREADS = 4,   // threads per block of the fast gather launches (0 = the shipped size; 128 / 256)
  SG_DEV_PEER_PUSH_BLOCKS = 5, // quarter-blocks per SM of the all-gather store kernel (0 = 1, i.e. one block per four SMs)
  SG_DEV_GEMM_TRACE = 8,        // 1: accumulate wait cycles of pair 0's leader CTA (sg_gemm_trace_read)
  SG_DEV_COUNT = 12
};
int dev_option(int which);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// ---- segment schedule ("plan") layout, shared by plan.cu and the gather kernels ----
// One work item = a run of at most `chunk` consecutive edges of one segment.
//   x: first edge, y: one-past-last edge, z: segment id, w: partial slot (-1: writes the output)
struct PlanHeader {
  int n_items;     // written by the build kernels
  int n_long;      // segments cut into > 1 item
  int n_partials;  // partial rows in use
  int chunk;
  int n_seg;
  int nnz;
  int cap_items;
  int cap_long;
};
// {segment, first partial slot, number of partials, unused}
struct PlanView {
  const PlanHeader *hdr;
  const int4 *items;
  const int4 *longs;
};

inline size_t plan_cap_items(int n_seg, int nnz, int chunk) { return (size_t)n_seg + (size_t)nnz / chunk + 1; }
inline size_t plan_cap_long(int nnz, int chunk) { return (size_t)nnz / chunk + 1; }
inline size_t plan_off_items() { return 64; }
inline size_t plan_off_longs(int n_seg, int nnz, int chunk) {
  return plan_off_items() + align_up(plan_cap_items(n_seg, nnz, chunk) * sizeof(int4), 64);
}
inline size_t plan_off_scan(int n_seg, int nnz, int chunk) {
  return plan_off_longs(n_seg, nnz, chunk) + align_up(plan_cap_long(nnz, chunk) * sizeof(int4), 64);
}

// exclusive scan of n int32 (out may alias in); ws needs scan_ws_bytes(n)
size_t scan_ws_bytes(int n);
int exclusive_scan_i32(int32_t *out, const int32_t *in, int n, int32_t *total /*device, may be null*/,
                       void *ws, cudaStream_t stream);

}  // namespace sg
