// Multi-GPU exchange of the node-partitioned aggregation over NVLink 5 / NVSwitch PEER MEMORY.
//
// The reference is single-device (experiments/STAR-GCN.py:32); this file is the B200-native transport of the
// dense-halo case (every rank needs nearly every remote neighbour row — a rating graph has no locality).  Every
// rank maps every other rank's exchange buffer into its address space (symmetric memory: the host side hands
// this library one device pointer per rank) and the collectives become plain stores, one flag barrier and a
// local fixed-order sum:
//
//   all-gather   (forward)   peer_push_rows_kernel: a rank reads its block of neighbour rows ONCE and stores it
//                            into the [n_total, D] table of every rank (its own included)
//   reduce-scatter (backward) fused into the PRODUCER: the transposed gather (gather.cu, PEER epilogue) stores
//                            each finished gradient row straight into the owner's staging slot of this rank, so
//                            the NVLink transfer rides inside the gather launch; after the barrier the owner sums
//                            its `world` slots in rank order (peer_reduce_kernel) — no float atomics, bit-identical
//                            reruns, same value on every run regardless of arrival order
//   all-reduce   (dW)        push to every rank's slot + the same local sum
//   barrier                  peer_barrier_kernel: release-store of an epoch into every peer's flag word, acquire-spin
//                            on the own flag words; the epoch lives in device memory so that the launch is
//                            CUDA-graph replayable.  A spin that exceeds `timeout_ns` records the missing rank in
//                            state[1] and returns (the caller checks it) — it never hangs the device.
//
// Hazards: one exchange alternates  push -> barrier(F) -> consume  and  scatter -> barrier(B) -> reduce  on ONE
// stream.  A rank can only overwrite a peer's table for step k+1 after it passed barrier B of step k, which the
// peer arrives at after its step-k consumers (stream order); a rank can only scatter into a peer's staging for
// step k+1 after barrier F of step k+1, which the peer arrives at after its step-k reduce.  Single buffers suffice.
#include "common.cuh"

namespace sg {

constexpr int kMaxPeers = SG_MAX_PEERS;

struct PeerRows { float4 *p[kMaxPeers]; };
struct PeerFlags { uint32_t *p[kMaxPeers]; };

__global__ void __launch_bounds__(256) peer_push_rows_kernel(const PeerRows dst, const float4 *__restrict__ src,
                                                             long long n4, int world) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {   // four independent loads in flight, then 4 x world stores
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < world) {
#pragma unroll
        for (int u = 0; u < 4; ++u) dst.p[q][i + u * stride] = v[u];
      }
  }
  for (; i < n4; i += stride) {
    const float4 v = __ldg(src + i);
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < world) dst.p[q][i] = v;
  }
  __threadfence_system();
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// state[0]: epoch of the last completed barrier; state[1]: 0, or 1 + the first rank that did not arrive in time
__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerFlags flags, uint32_t *state, int rank, int world,
                                                          unsigned long long timeout_ns) {
  const uint32_t epoch = state[0] + 1;   // every lane reads the same word before lane 0 rewrites it below
  const uint32_t failed = state[1];
  __syncwarp();
  if (failed) return;                    // a barrier of this exchange already timed out: do not wait again
  const int q = threadIdx.x;
  if (q < world) {
    uint32_t *theirs = nullptr, *mine = nullptr;
#pragma unroll
    for (int s = 0; s < kMaxPeers; ++s) {
      if (s == q) theirs = flags.p[s] + rank;     // my arrival, in rank q's flag array
      if (s == rank) mine = flags.p[s] + q;       // rank q's arrival, in my flag array
    }
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const unsigned long long t0 = global_ns();
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      if (global_ns() - t0 > timeout_ns) {
        atomicCAS(state + 1, 0u, (uint32_t)(q + 1));
        break;
      }
      __nanosleep(64);
    }
  }
  __syncwarp();
  if (threadIdx.x == 0) state[0] = epoch;
}

// out[i] (=|+=) stage[0][i] + stage[1][i] + ... in rank order
__global__ void __launch_bounds__(256) peer_reduce_kernel(float4 *__restrict__ out, const float4 *__restrict__ stage,
                                                          long long n4, long long slot4, int world, int add) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < world) v[q] = stage[q * slot4 + i];
    float4 acc = v[0];
#pragma unroll
    for (int q = 1; q < kMaxPeers; ++q)
      if (q < world) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
    if (add) {
      const float4 o = out[i];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    out[i] = acc;
  }
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace sg

using namespace sg;

extern "C" {

int sg_peer_push_rows(float *const *dst_host, const float *src, long long n_floats, int world, sg_stream_t stream) {
  SG_REQUIRE(world >= 1 && world <= kMaxPeers, "sg_peer_push_rows: world must be 1..%d", kMaxPeers);
  SG_REQUIRE(n_floats >= 0 && n_floats % 4 == 0, "sg_peer_push_rows: n_floats must be a multiple of 4");
  if (n_floats == 0) return SG_OK;
  SG_REQUIRE(dst_host && src && aligned16(src), "sg_peer_push_rows: null or unaligned source");
  PeerRows d{};
  for (int q = 0; q < world; ++q) {
    SG_REQUIRE(dst_host[q] && aligned16(dst_host[q]), "sg_peer_push_rows: destination %d null or not 16-byte aligned", q);
    d.p[q] = reinterpret_cast<float4 *>(dst_host[q]);
  }
  const long long n4 = n_floats / 4;
  long long blocks = ceil_div<long long>(n4, 256 * 4);
  const int dv = dev_option(SG_DEV_PEER_PUSH_BLOCKS);
  // The stores are posted and the transfer is NVLink-bound (0.20 ms for a 17.9 MB block to 8 tables whatever the
  // grid), so a SMALL grid is enough — and it leaves the SMs to the other layer direction's gather, which runs at the
  // same time.  8 GPUs, weak scaling, step time by grid: 296 CTAs 1.794 ms, 148: 1.790, 74: 1.750, 37: 1.690
  // (profiles/r02_summary.md I).  One block per four SMs.
  const long long cap = world == 1 ? 4LL * num_sms()          // a local copy (the sparse layout's own rows): HBM-bound
                                   : (dv > 0 ? dv : 1) * (long long)num_sms() / 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  peer_push_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d, reinterpret_cast<const float4 *>(src), n4, world);
  SG_LAUNCHED("peer_push_rows_kernel");
  return SG_OK;
}

int sg_peer_barrier(uint32_t *const *flags_host, uint32_t *state, int rank, int world, double timeout_s,
                    sg_stream_t stream) {
  SG_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "sg_peer_barrier: bad rank / world");
  SG_REQUIRE(flags_host && state && timeout_s > 0, "sg_peer_barrier: null pointer or bad timeout");
  PeerFlags f{};
  for (int q = 0; q < world; ++q) {
    SG_REQUIRE(flags_host[q], "sg_peer_barrier: flag array of rank %d is null", q);
    f.p[q] = flags_host[q];
  }
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, state, rank, world, (unsigned long long)(timeout_s * 1e9));
  SG_LAUNCHED("peer_barrier_kernel");
  return SG_OK;
}

int sg_peer_reduce(float *out, const float *stage, long long n_floats, long long slot_stride_floats, int world, int req,
                   sg_stream_t stream) {
  SG_REQUIRE(valid_req(req), "sg_peer_reduce: bad req %d", req);
  if (req == SG_REQ_NULL || n_floats == 0) return SG_OK;
  SG_REQUIRE(world >= 1 && world <= kMaxPeers, "sg_peer_reduce: world must be 1..%d", kMaxPeers);
  SG_REQUIRE(n_floats > 0 && n_floats % 4 == 0 && slot_stride_floats % 4 == 0 && slot_stride_floats >= n_floats,
             "sg_peer_reduce: sizes must be multiples of 4 floats and slots must not overlap");
  SG_REQUIRE(out && stage && aligned16(out) && aligned16(stage), "sg_peer_reduce: null or unaligned pointer");
  const long long n4 = n_floats / 4;
  long long blocks = ceil_div<long long>(n4, 256);
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  peer_reduce_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4 *>(out), reinterpret_cast<const float4 *>(stage), n4, slot_stride_floats / 4, world,
      req == SG_REQ_ADD ? 1 : 0);
  SG_LAUNCHED("peer_reduce_kernel");
  return SG_OK;
}

}  // extern "C"
