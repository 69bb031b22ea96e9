// Relation transform on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate.
//
// The per-rating-level FullyConnected + add_n of MultiLinkGCNAggregator.hybrid_forward
// (mxgraph/layers/aggregators.py:141-159) is, after aggregate-first reordering, ONE dense GEMM
//     Z[n_dst, U] = [agg | wsum][n_dst, R*D + R] . [W_cat | B^T][U, R*D + R]^T
// and its backward two more (dAgg = gZ . W_cat,  dW = gZ^T . [agg | wsum]).  fp32 parity at 1e-5
// rules out plain TF32 (2^-11 input rounding), so every operand arrives pre-split as
//     x = x_hi + x_lo,   x_hi = x with the low 13 mantissa bits cleared (exact in TF32)
// and the kernel accumulates  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  into one fp32 TMEM accumulator
// (the dropped A_lo.B_lo term is ~2^-22 relative).
//
// Kernel shape (persistent; a CTA PAIR — cluster of 2, tcgen05 cta_group::2 — owns one 256 x 256 output tile):
//   TMA warp    cp.async.bulk.tensor (SWIZZLE_128B) -> 3-deep smem ring, each CTA ITS 128 rows of A and ITS half of B
//   MMA warp    (leader CTA) one elected lane issues tcgen05.mma.cta_group::2.kind::tf32 (M=256, N<=256, K=8) from
//               smem descriptors; tcgen05.commit (multicast) frees the stage in both CTAs
//   8 epilogue warps  drain each TMEM chain (tcgen05.ld 32x32b.x32) into fp32 register accumulators; at the end
//               of a tile bias + activation and coalesced stores through a padded smem transpose, while the MMA
//               warp already runs the next tile's chains
//   4 splitter warps (tf32x3_gemm_split_kernel) do the hi/lo split of operands that arrive as plain fp32
// Two kernels: tf32x3_gemm_pair_kernel takes operands pre-split by a producer pass (small Dense layers, weights);
// tf32x3_gemm_split_kernel splits inside the kernel (the large activations: agg, gZ).
// Operand layouts: K-major (A[M,K], B[N,K] row-major; forward and dAgg) or MN-major (A[K,M],
// B[K,N] row-major; the weight gradient, whose reduction axis is the node axis) — the latter
// loads [32 floats x 32 k-rows] swizzle atoms (one TMA box each) and sets the a_major/b_major
// bits of the instruction descriptor.  Split-K partials are summed in a fixed order.
#include <cuda.h>

#include "common.cuh"

namespace sg {

constexpr int kBM = 128;        // UMMA M
constexpr int kBK = 32;         // floats per stage row = 128 B = one swizzle span
constexpr int kUmmaK = 8;       // tf32 MMA K
constexpr int kGemmThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps

struct GemmArgs {
  float *D;
  long long split_stride;  // elements between split-K partial outputs
  int ldd;
  int M, N;
  int kb_total;      // number of BK-wide k-blocks
  int kb_per_split;
  int tiles_m, tiles_n, splits;
  int chain_kb;      // k-blocks per TMEM accumulation chain (kChainKBlocks)
  int epi;           // 0 store, 1 leaky (slope)
  float slope;
  const float *bias; // optional [N], added before the activation
  int trace;          // development: accumulate wait cycles of pair 0's leader into g_gemm_trace
  int relaxed_arrive; // hand the TMEM buffer back with a relaxed arrival (default; see mbar_arrive_leader_relaxed)
};

// Development trace (SG_DEV_GEMM_TRACE): cycles the roles of the LEADER CTA of pair 0 spend waiting, summed over the
// launch — [0] MMA thread on full_bar, [1] MMA thread on tmem_empty, [2] MMA thread total, [3] producer 0 on
// empty_bar, [4] producer 0 total, [5] epilogue warp on tmem_full, [6] epilogue warp total, [7] k-blocks.
__device__ unsigned long long g_gemm_trace[8];
struct TraceClock {
  long long t0;
  bool on;
  __device__ __forceinline__ TraceClock(bool enabled) : t0(0), on(enabled) {}
  __device__ __forceinline__ void begin() { if (on) t0 = clock64(); }
  __device__ __forceinline__ void end(int slot) { if (on) atomicAdd(&g_gemm_trace[slot], (unsigned long long)(clock64() - t0)); }
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride
// byte offsets (all >> 4), version 1 (Blackwell), layout type 2 = SWIZZLE_128B (16-byte chunks over 8 rows);
// layout 1 = SWIZZLE_128B_BASE32B (32-byte chunks permuted over 4 rows): the only layout the hardware
// accepts for MN-major 32-bit operands (cutlass sm100_common.inl:92).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}

// ---------------------------------------------------------------------------------------------
// the GEMM kernel
//
// Accumulation accuracy: the tensor core adds into its fp32 TMEM accumulator with truncation
// (measured: a 252-MMA chain at K=650 lands 3.5e-6 relative BELOW the exact magnitude, always
// toward zero; about 2e-8 relative per accumulating MMA).  To stay well inside the 1e-5 parity bar at
// any K, a TMEM accumulator only ever holds a chain of kChainKBlocks k-blocks (48 MMAs, <= 1e-6); the
// epilogue warps drain it (tcgen05.ld) and add it into fp32 REGISTER accumulators with round-to-nearest
// while the MMA warp fills the other TMEM buffer (2 x BN columns = all 512 TMEM columns).  The chain
// length trades time for accuracy — forward transform / error of one fused layer vs the fp64 answer:
//   chain 2: 0.160 ms / 7.7e-7    chain 4: 0.136 ms / 8.8e-7    chain 8: 0.122 ms / 1.3e-6
// (a plain fp32 FMA GEMM is at 1e-7 .. 3e-7).
// ---------------------------------------------------------------------------------------------
constexpr int kChainKBlocks = 4;  // 4 k-blocks x 3 hi/lo products x 4 = 48 MMAs per TMEM chain
constexpr int kEpiWarps = 8;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster own one 256 x 256 output tile.
// Each CTA loads ITS 128 rows of A and ITS half of the B tile (hi + lo: 64 KB per k-block instead of
// 96 KB), the leader CTA issues M=256 MMAs that read both CTAs' shared memory and write both CTAs'
// TMEM, and tcgen05.commit multicasts the stage-free / accumulator-ready arrivals to both CTAs.
// Why: with fp32 operands pre-split into hi + lo the kernel streams 8 bytes per operand element, and the
// B tile (the weights) is re-streamed from L2 by every M tile — the single-CTA kernel moves 1.1 GB
// through the L2->SM fabric for the ML-10M forward transform and is bound by it, not by the tensor pipe
// (issuing one of the three products instead of all three changes its time by 15 %).  A pair halves the
// B traffic per SM (64 KB stages, 3-deep ring).  Measured: 0.647 -> 0.545 ms for the six GEMMs of a step.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's even CTA

__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint64_t *leader_bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {  // arrive on the leader CTA's copy of `bar`
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)) : "memory");
}
// Same arrival without the cluster-scope release.  The release makes the warp wait for every store it has in
// flight — the 128 output-row stores of the tile it has just written — before the TMEM buffer is handed back
// (ncu source page, final capture of round 1: 15-26 % of all warp samples of the pair kernel sit on the
// ERRBAR / SYNCS.ARRIVE pair of this arrival with stall_membar).  The hand-over only has to order the
// tcgen05.ld reads, which tcgen05.wait::ld (the values are in registers before the arrival is issued) +
// tcgen05.fence::before_thread_sync already do.  Default since round 2 (full GPU suite green).
__device__ __forceinline__ void mbar_arrive_leader_relaxed(uint64_t *bar) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)) : "memory");
}

template <int STAGES, bool MN_MAJOR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
    tf32x3_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                            const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                            const GemmArgs g) {
  constexpr int BN = 256;                  // tile columns (UMMA N); each CTA stores all BN columns of its 128 rows
  constexpr int A_BYTES = kBM * kBK * 4;   // this CTA's 128 rows of A
  constexpr int B_BYTES = 128 * kBK * 4;   // this CTA's half of the B tile
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr int STG_FLOATS = 32 * 33;
  constexpr uint32_t IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | ((MN_MAJOR ? 1u : 0u) << 15) |
                                  ((MN_MAJOR ? 1u : 0u) << 16) | ((uint32_t)(256 >> 4) << 24);  // M = 256 over the pair

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float *stg_all = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES);
  __shared__ uint64_t full_bar[STAGES];    // used on the leader only (both CTAs' TMA bytes land on it)
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];   // used on the leader only (both CTAs' epilogue warps arrive)
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_n = g.tiles_n, tiles_mn = g.tiles_m * g.tiles_n;  // tiles_m counts 256-row pair tiles
  const int n_tiles = tiles_mn * g.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_b_hi); tma_prefetch_desc(&map_b_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 2 * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  TraceClock trc(g.trace && pair_id == 0 && leader && lane == 0 && (warp == 0 || warp == 1 || warp == 2));
  const long long trc_start = trc.on ? clock64() : 0;

  if (warp == 0) {
    if (lane == 0) {
      int v = 0;
      for (int tile = pair_id; tile < n_tiles; tile += n_pairs) {
        const int z = tile / tiles_mn, rem = tile - z * tiles_mn;
        const int m0 = (rem / tiles_n) * 256 + (int)rank * kBM, n0 = (rem % tiles_n) * BN;
        const int n_eff = min(BN, ((g.N - n0) + 15) & ~15);
        const int nb0 = n0 + (int)rank * (n_eff >> 1);  // this CTA supplies B rows [nb0, nb0 + n_eff/2)
        const int kb_begin = z * g.kb_per_split, kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++v) {
          const int s = v % STAGES;
          const uint32_t ph = (v / STAGES) & 1;
          trc.begin();
          mbar_wait(&empty_bar[s], ph ^ 1);
          trc.end(3);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);  // bytes of BOTH CTAs
          uint8_t *sa_hi = smem + s * STAGE_BYTES, *sa_lo = sa_hi + A_BYTES, *sb_hi = sa_lo + A_BYTES, *sb_lo = sb_hi + B_BYTES;
          if constexpr (MN_MAJOR) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              tma_load_2d_pair(sa_hi + a * (kBK * 128), &map_a_hi, &full_bar[s], m0 + a * 32, kb * kBK);
              tma_load_2d_pair(sa_lo + a * (kBK * 128), &map_a_lo, &full_bar[s], m0 + a * 32, kb * kBK);
              tma_load_2d_pair(sb_hi + a * (kBK * 128), &map_b_hi, &full_bar[s], nb0 + a * 32, kb * kBK);
              tma_load_2d_pair(sb_lo + a * (kBK * 128), &map_b_lo, &full_bar[s], nb0 + a * 32, kb * kBK);
            }
          } else {
            tma_load_2d_pair(sa_hi, &map_a_hi, &full_bar[s], kb * kBK, m0);
            tma_load_2d_pair(sa_lo, &map_a_lo, &full_bar[s], kb * kBK, m0);
            tma_load_2d_pair(sb_hi, &map_b_hi, &full_bar[s], kb * kBK, nb0);
            tma_load_2d_pair(sb_lo, &map_b_lo, &full_bar[s], kb * kBK, nb0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      int v = 0, chain = 0;
      for (int tile = pair_id; tile < n_tiles; tile += n_pairs) {
        const int z = tile / tiles_mn, rem = tile - z * tiles_mn;
        const int n0 = (rem % tiles_n) * BN;
        const int kb_begin = z * g.kb_per_split, kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
        const int n_kb = kb_end - kb_begin;
        const int n_eff = min(BN, ((g.N - n0) + 15) & ~15);
        const uint32_t idesc = IDESC_BASE | ((uint32_t)(n_eff >> 3) << 17);
        // chains of a tile are balanced (21 k-blocks -> 4+4+4+3+3+3, not 4x5+1): a one-k-block tail chain finishes
        // before the buffer of the chain before the previous one has been drained and stalls the MMA warp
        const int n_chains = (n_kb + g.chain_kb - 1) / g.chain_kb;
        const int len_lo = n_kb / n_chains, n_long = n_kb - len_lo * n_chains;   // the first n_long chains have len_lo + 1
        int vin = 0, clen = len_lo + (n_long > 0 ? 1 : 0), cidx = 0;
        for (int i = 0; i < n_kb; ++i, ++v) {
          const int s = v % STAGES;
          const uint32_t ph = (v / STAGES) & 1;
          const int buf = chain & 1;
          if (vin == 0) {
            trc.begin();
            mbar_wait(&tmem_empty_bar[buf], ((chain >> 1) & 1) ^ 1);
            trc.end(1);
            tc_fence_after();
          }
          trc.begin();
          mbar_wait(&full_bar[s], ph);
          trc.end(0);
          tc_fence_after();
          const uint32_t sa_hi = smem_u32(smem + s * STAGE_BYTES), sa_lo = sa_hi + A_BYTES, sb_hi = sa_lo + A_BYTES,
                         sb_lo = sb_hi + B_BYTES;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {  // hi.hi, hi.lo, then lo.hi (the order the split kernel needs)
            const uint32_t sa = pass == 2 ? sa_lo : sa_hi, sb = pass == 1 ? sb_lo : sb_hi;
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              uint64_t da, db;
              if constexpr (MN_MAJOR) {
                da = make_smem_desc(sa + k * 1024, kBK * 128, 512, 1);
                db = make_smem_desc(sb + k * 1024, kBK * 128, 512, 1);
              } else {
                da = make_smem_desc(sa + k * 32, 16, 1024, 2);
                db = make_smem_desc(sb + k * 32, 16, 1024, 2);
              }
              umma_tf32_pair(tmem_base + (uint32_t)(buf * BN), da, db, idesc, (vin != 0) || (pass != 0) || (k != 0));
            }
          }
          umma_commit_pair(&empty_bar[s]);
          if (trc.on) atomicAdd(&g_gemm_trace[7], 1ull);
          if (++vin == clen) {
            umma_commit_pair(&tmem_full_bar[buf]);
            ++chain;
            ++cidx;
            vin = 0;
            clen = len_lo + (cidx < n_long ? 1 : 0);
          }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int cbase = h * (BN / 2);
    float *stg = stg_all + (warp - 2) * STG_FLOATS;
    int chain = 0;
    for (int tile = pair_id; tile < n_tiles; tile += n_pairs) {
      const int z = tile / tiles_mn, rem = tile - z * tiles_mn;
      const int m0 = (rem / tiles_n) * 256 + (int)rank * kBM, n0 = (rem % tiles_n) * BN;
      const int kb_begin = z * g.kb_per_split, kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
      const int n_chains = (kb_end - kb_begin + g.chain_kb - 1) / g.chain_kb;
      const bool active = n0 + cbase < g.N && m0 < g.M;  // warp-uniform
      float racc[BN / 2];
#pragma unroll
      for (int j = 0; j < BN / 2; ++j) racc[j] = 0.f;
      for (int c = 0; c < n_chains; ++c, ++chain) {
        const int buf = chain & 1;
        trc.begin();
        mbar_wait(&tmem_full_bar[buf], (chain >> 1) & 1);
        trc.end(5);
        tc_fence_after();
        if (active) {
#pragma unroll
          for (int ch = 0; ch < BN / 2 / 32; ++ch) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cbase + ch * 32), r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) racc[ch * 32 + j] += __uint_as_float(r[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (g.relaxed_arrive) mbar_arrive_leader_relaxed(&tmem_empty_bar[buf]);
          else mbar_arrive_leader(&tmem_empty_bar[buf]);
        }
      }
      if (active) {
        float *dbase = g.D + (long long)z * g.split_stride;
        const int row0 = m0 + q * 32;
#pragma unroll
        for (int ch = 0; ch < BN / 2 / 32; ++ch) {
          const int col = n0 + cbase + ch * 32 + lane;
          if (n0 + cbase + ch * 32 >= g.N) break;
          float bias = 0.f;
          if (g.bias && col < g.N) bias = __ldg(g.bias + col);
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = racc[ch * 32 + j];
          __syncwarp();
          const int rows = min(32, g.M - row0);
          const bool col_ok = col < g.N;
          float *dcol = dbase + (long long)row0 * g.ldd + col;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {      // unrolled: 8 independent smem reads / row stores in flight
            float x = stg[rr * 33 + lane] + bias;
            if (g.epi == 1) x = x > 0.f ? x : g.slope * x;
            if (col_ok && rr < rows) dcol[(long long)rr * g.ldd] = x;
          }
          __syncwarp();
        }
      }
    }
  }
  if (trc.on) atomicAdd(&g_gemm_trace[warp == 1 ? 2 : (warp == 0 ? 4 : 6)], (unsigned long long)(clock64() - trc_start));
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------
// CTA-pair kernel with the hi/lo operand split done INSIDE the kernel.
//
// The pre-split kernel above streams 8 bytes per operand element (hi + lo) through the L2->SM fabric and needs a
// producer pass that writes both halves to HBM.  Here an operand arrives as plain fp32.  Two facts make that cheap:
//   * the tensor core TRUNCATES fp32 inputs to TF32 (measured: feeding the raw tile as the "hi" operand gives
//     bit-identical results to feeding x & 0xffffe000, tools/gemm_bench.py) — so the raw tile TMA delivers IS the
//     hi operand and only the remainder lo = x - trunc(x) has to be produced;
//   * of the three products  hi.hi + hi.lo + lo.hi  only those with a "lo" factor wait for that remainder.
// Four splitter warps compute lo from the raw tile (element-wise, so the swizzled layout TMA produced is preserved),
// `fence.proxy.async` makes their generic-proxy stores visible to the tensor core's async-proxy reads, and a second
// barrier on the leader CTA (4 warps x 2 CTAs arrivals) releases the lo products; the MMA warp issues the hi
// products as soon as the TMA bytes of BOTH CTAs have landed (same critical path as the pre-split kernel, 25 - 50 %
// fewer bytes).  All TMA loads complete on the LEADER's full barrier (cta_group::2); the leader's splitter relays
// that to the peer CTA's splitter warps through `go_bar`.  A is always raw; B is pre-split (small weight matrices,
// split once per call) or raw (SPLIT_B: the weight gradient, whose B operand is the aggregated feature matrix).
//
// 512 threads = 4 warpgroups with their own register budgets (setmaxnreg): WG0 = TMA warp, MMA warp, TMEM
// allocator (40 registers), WG1-2 = 8 epilogue warps holding the 128 fp32 accumulators of the chained TMEM
// drain (208), WG3 = 4 splitter warps (56); 128 x (40 + 208 + 208 + 56) = the whole register file.
// ---------------------------------------------------------------------------------------------
constexpr int kSplitThreads = 512;
constexpr int kSplitterWarps = 4;

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// lo = x - trunc_tf32(x) of one 16 KB operand tile by the 128 splitter threads (conflict-free 16-byte accesses);
// the raw tile stays in place as the hi operand
__device__ __forceinline__ void split_tile_16k(const uint8_t *raw_slot, uint8_t *lo_slot, int st) {
  constexpr int kVecs = 16384 / 16 / (kSplitterWarps * 32);  // 8 float4 per thread
  float4 v[kVecs];
#pragma unroll
  for (int i = 0; i < kVecs; ++i) v[i] = reinterpret_cast<const float4 *>(raw_slot)[i * (kSplitterWarps * 32) + st];
#pragma unroll
  for (int i = 0; i < kVecs; ++i) {
    float4 l;
    l.x = v[i].x - __uint_as_float(__float_as_uint(v[i].x) & 0xffffe000u);
    l.y = v[i].y - __uint_as_float(__float_as_uint(v[i].y) & 0xffffe000u);
    l.z = v[i].z - __uint_as_float(__float_as_uint(v[i].z) & 0xffffe000u);
    l.w = v[i].w - __uint_as_float(__float_as_uint(v[i].w) & 0xffffe000u);
    reinterpret_cast<float4 *>(lo_slot)[i * (kSplitterWarps * 32) + st] = l;
  }
}

__device__ __forceinline__ void mbar_arrive_peer(uint64_t *bar) {  // arrive on the non-leader CTA's copy of `bar`
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)) : "memory");
}

template <int STAGES, bool MN_MAJOR, bool SPLIT_B>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSplitThreads, 1)
    tf32x3_gemm_split_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_hi,
                             const __grid_constant__ CUtensorMap map_b_lo, const GemmArgs g) {
  constexpr int BN = 256;
  constexpr int A_BYTES = kBM * kBK * 4;   // this CTA's 128 rows of A (one slot; hi and lo slots per stage)
  constexpr int B_BYTES = 128 * kBK * 4;   // this CTA's half of the B tile
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr int TX_BYTES = A_BYTES + (SPLIT_B ? B_BYTES : 2 * B_BYTES);   // what TMA delivers per stage and CTA
  constexpr int STG_FLOATS = 32 * 33;
  constexpr uint32_t IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | ((MN_MAJOR ? 1u : 0u) << 15) |
                                  ((MN_MAJOR ? 1u : 0u) << 16) | ((uint32_t)(256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float *stg_all = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES);
  __shared__ uint64_t full_bar[STAGES];    // leader only: TMA bytes of BOTH CTAs
  __shared__ uint64_t go_bar[STAGES];      // peer only: the leader saw full_bar complete (relayed by its splitter)
  __shared__ uint64_t empty_bar[STAGES];   // MMAs that read the stage have completed (multicast commit)
  __shared__ uint64_t split_bar[STAGES];   // leader only: lo tiles written in both CTAs (2 x kSplitterWarps arrivals)
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];   // leader only
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_n = g.tiles_n, tiles_mn = g.tiles_m * g.tiles_n;
  const int n_tiles = tiles_mn * g.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b_hi);
    if constexpr (!SPLIT_B) tma_prefetch_desc(&map_b_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1); mbar_init(&go_bar[s], 1); mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], 2 * kSplitterWarps);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 2 * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp < 4) {
    reg_dec<40>();
    if (warp == 0 && lane == 0) {
      // ---- TMA producer (both CTAs): raw A tile -> hi slot; B pre-split (hi, lo) or raw -> hi slot ----
      int v = 0;
      for (int tile = pair_id; tile < n_tiles; tile += n_pairs) {
        const int z = tile / tiles_mn, rem = tile - z * tiles_mn;
        const int m0 = (rem / tiles_n) * 256 + (int)rank * kBM, n0 = (rem % tiles_n) * BN;
        const int n_eff = min(BN, ((g.N - n0) + 15) & ~15);
        const int nb0 = n0 + (int)rank * (n_eff >> 1);
        const int kb_begin = z * g.kb_per_split, kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++v) {
          const int s = v % STAGES;
          const uint32_t ph = (v / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * TX_BYTES);  // bytes of BOTH CTAs land on the leader's barrier
          uint8_t *sa_hi = smem + s * STAGE_BYTES, *sb_hi = sa_hi + 2 * A_BYTES, *sb_lo = sb_hi + B_BYTES;
          if constexpr (MN_MAJOR) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              tma_load_2d_pair(sa_hi + a * (kBK * 128), &map_a, &full_bar[s], m0 + a * 32, kb * kBK);
              tma_load_2d_pair(sb_hi + a * (kBK * 128), &map_b_hi, &full_bar[s], nb0 + a * 32, kb * kBK);
              if constexpr (!SPLIT_B) tma_load_2d_pair(sb_lo + a * (kBK * 128), &map_b_lo, &full_bar[s], nb0 + a * 32, kb * kBK);
            }
          } else {
            tma_load_2d_pair(sa_hi, &map_a, &full_bar[s], kb * kBK, m0);
            tma_load_2d_pair(sb_hi, &map_b_hi, &full_bar[s], kb * kBK, nb0);
            if constexpr (!SPLIT_B) tma_load_2d_pair(sb_lo, &map_b_lo, &full_bar[s], kb * kBK, nb0);
          }
        }
      }
    } else if (warp == 1 && leader && lane == 0) {
      // ---- MMA issuer (leader CTA): hi products when the TMA bytes have landed, lo products after the split ----
      int v = 0, chain = 0;
      for (int tile = pair_id; tile < n_tiles; tile += n_pairs) {
        const int z = tile / tiles_mn, rem = tile - z * tiles_mn;
        const int n0 = (rem % tiles_n) * BN;
        const int kb_begin = z * g.kb_per_split, kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
        const int n_kb = kb_end - kb_begin;
        const int n_eff = min(BN, ((g.N - n0) + 15) & ~15);
        const uint32_t idesc = IDESC_BASE | ((uint32_t)(n_eff >> 3) << 17);
        // chains of a tile are balanced (21 k-blocks -> 4+4+4+3+3+3, not 4x5+1): a one-k-block tail chain finishes
        // before the buffer of the chain before the previous one has been drained and stalls the MMA warp
        const int n_chains = (n_kb + g.chain_kb - 1) / g.chain_kb;
        const int len_lo = n_kb / n_chains, n_long = n_kb - len_lo * n_chains;   // the first n_long chains have len_lo + 1
        int vin = 0, clen = len_lo + (n_long > 0 ? 1 : 0), cidx = 0;
        for (int i = 0; i < n_kb; ++i, ++v) {
          const int s = v % STAGES;
          const uint32_t ph = (v / STAGES) & 1;
          const int buf = chain & 1;
          if (vin == 0) {
            mbar_wait_cluster(&tmem_empty_bar[buf], ((chain >> 1) & 1) ^ 1);
            tc_fence_after();
          }
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa_hi = smem_u32(smem + s * STAGE_BYTES), sa_lo = sa_hi + A_BYTES, sb_hi = sa_lo + A_BYTES,
                         sb_lo = sb_hi + B_BYTES;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {  // hi.hi, hi.lo, lo.hi
            // the raw tiles are the hi operands; a pre-split B_lo arrives by TMA, everything else "lo" by the splitters
            if ((pass == 1 && SPLIT_B) || (pass == 2 && !SPLIT_B)) {
              mbar_wait_cluster(&split_bar[s], ph);
              tc_fence_after();
            }
            const uint32_t sa = pass == 2 ? sa_lo : sa_hi, sb = pass == 1 ? sb_lo : sb_hi;
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              uint64_t da, db;
              if constexpr (MN_MAJOR) {
                da = make_smem_desc(sa + k * 1024, kBK * 128, 512, 1);
                db = make_smem_desc(sb + k * 1024, kBK * 128, 512, 1);
              } else {
                da = make_smem_desc(sa + k * 32, 16, 1024, 2);
                db = make_smem_desc(sb + k * 32, 16, 1024, 2);
              }
              umma_tf32_pair(tmem_base + (uint32_t)(buf * BN), da, db, idesc, (vin != 0) || (pass != 0) || (k != 0));
            }
          }
          umma_commit_pair(&empty_bar[s]);
          if (++vin == clen) {
            umma_commit_pair(&tmem_full_bar[buf]);
            ++chain;
            ++cidx;
            vin = 0;
            clen = len_lo + (cidx < n_long ? 1 : 0);
          }
        }
      }
    }
  } else if (warp >= 12) {
    // ---- splitter warps (both CTAs): lo = x - trunc(x) of the raw tiles ----
    reg_dec<56>();
    const int st = threadIdx.x - 12 * 32;
    int v = 0;
    for (int tile = pair_id; tile < n_tiles; tile += n_pairs) {
      const int z = tile / tiles_mn;
      const int kb_begin = z * g.kb_per_split, kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
      for (int kb = kb_begin; kb < kb_end; ++kb, ++v) {
        const int s = v % STAGES;
        const uint32_t ph = (v / STAGES) & 1;
        if (leader) {
          mbar_wait(&full_bar[s], ph);
          if (warp == 12 && lane == 0) mbar_arrive_peer(&go_bar[s]);   // relay: the peer's bytes have landed as well
        } else {
          mbar_wait_cluster(&go_bar[s], ph);
        }
        uint8_t *sa_hi = smem + s * STAGE_BYTES;
        split_tile_16k(sa_hi, sa_hi + A_BYTES, st);
        if constexpr (SPLIT_B) split_tile_16k(sa_hi + 2 * A_BYTES, sa_hi + 2 * A_BYTES + B_BYTES, st);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&split_bar[s]);
      }
    }
  } else {
    // ---- epilogue warps 4..11: TMEM lane quarter q = warp % 4, column half h ----
    reg_inc<208>();
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    const int cbase = h * (BN / 2);
    float *stg = stg_all + (warp - 4) * STG_FLOATS;
    int chain = 0;
    for (int tile = pair_id; tile < n_tiles; tile += n_pairs) {
      const int z = tile / tiles_mn, rem = tile - z * tiles_mn;
      const int m0 = (rem / tiles_n) * 256 + (int)rank * kBM, n0 = (rem % tiles_n) * BN;
      const int kb_begin = z * g.kb_per_split, kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
      const int n_chains = (kb_end - kb_begin + g.chain_kb - 1) / g.chain_kb;
      const bool active = n0 + cbase < g.N && m0 < g.M;  // warp-uniform
      float racc[BN / 2];
#pragma unroll
      for (int j = 0; j < BN / 2; ++j) racc[j] = 0.f;
      for (int c = 0; c < n_chains; ++c, ++chain) {
        const int buf = chain & 1;
        mbar_wait(&tmem_full_bar[buf], (chain >> 1) & 1);
        tc_fence_after();
        if (active) {
#pragma unroll
          for (int ch = 0; ch < BN / 2 / 32; ++ch) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cbase + ch * 32), r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) racc[ch * 32 + j] += __uint_as_float(r[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (g.relaxed_arrive) mbar_arrive_leader_relaxed(&tmem_empty_bar[buf]);
          else mbar_arrive_leader(&tmem_empty_bar[buf]);
        }
      }
      if (active) {
        float *dbase = g.D + (long long)z * g.split_stride;
        const int row0 = m0 + q * 32;
#pragma unroll
        for (int ch = 0; ch < BN / 2 / 32; ++ch) {
          const int col = n0 + cbase + ch * 32 + lane;
          if (n0 + cbase + ch * 32 >= g.N) break;
          float bias = 0.f;
          if (g.bias && col < g.N) bias = __ldg(g.bias + col);
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = racc[ch * 32 + j];
          __syncwarp();
          const int rows = min(32, g.M - row0);
          const bool col_ok = col < g.N;
          float *dcol = dbase + (long long)row0 * g.ldd + col;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {      // unrolled: 8 independent smem reads / row stores in flight
            float x = stg[rr * 33 + lane] + bias;
            if (g.epi == 1) x = x > 0.f ? x : g.slope * x;
            if (col_ok && rr < rows) dcol[(long long)rr * g.ldd] = x;
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Measurement aid: what can TMA deliver to one SM?  Every CTA streams [32 floats x 128 rows] boxes (16 KB, the A
// tile of the GEMMs) of a row-major matrix through a ring of `stages` slots of `boxes` boxes each; a consumer
// thread frees a slot as soon as it has landed.  No MMA, no epilogue: GB/s per SM as a function of the bytes in
// flight separates a latency bound (rate grows with the ring) from a throughput bound (rate flat).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1) tma_probe_kernel(const __grid_constant__ CUtensorMap map, int rows, int kb_total,
                                                           int stages, int boxes, int iters, int same_boxes, int producers) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[16];
  __shared__ uint64_t empty_bar[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], producers); mbar_init(&empty_bar[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int tiles_m = rows / kBM;
  if (warp < producers && lane == 0) {       // producer warp p issues boxes p, p + producers, ... of every stage
    const int mine = (boxes - warp + producers - 1) / producers;
    for (int v = 0; v < iters; ++v) {
      const int s = v % stages;
      const uint32_t ph = (v / stages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      mbar_expect_tx(&full_bar[s], mine * kBM * kBK * 4);
      for (int b = warp; b < boxes; b += producers) {
        // a different box every time; same_boxes: every CTA walks the SAME boxes (hot L2 lines)
        const long long t = ((long long)(same_boxes ? 0 : blockIdx.x) * iters + v) * boxes + b;
        const int m0 = (int)(t % tiles_m) * kBM, kb = (int)((t / tiles_m) % kb_total);
        tma_load_2d(smem + (s * boxes + b) * (kBM * kBK * 4), &map, &full_bar[s], kb * kBK, m0);
      }
    }
  } else if (warp == 4 && lane == 0) {
    for (int v = 0; v < iters; ++v) {
      const int s = v % stages;
      const uint32_t ph = (v / stages) & 1;
      mbar_wait(&full_bar[s], ph);
      mbar_arrive(&empty_bar[s]);
    }
  }
}

// fixed-order sum of split-K partials
__global__ void __launch_bounds__(256) splitk_reduce_kernel(float *__restrict__ dst, int ldd, const float *__restrict__ ws,
                                                            int M, int N, int splits) {
  const long long total = (long long)M * N;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(long long)s * total + t];
    const int r = (int)(t / N), c = (int)(t - (long long)r * N);
    dst[(long long)r * ldd + c] = acc;
  }
}

// x -> (hi, lo): hi = x with the 13 low mantissa bits cleared (exactly representable in TF32),
// lo = x - hi (exact in fp32).  Optional transpose; destination padded to ld_dst with zeros.
__global__ void __launch_bounds__(256) split_tf32_kernel(float *__restrict__ hi, float *__restrict__ lo, int ld_dst,
                                                         const float *__restrict__ src, int rows, int cols, int ld_src,
                                                         int transpose) {
  const int out_rows = transpose ? cols : rows, out_cols = transpose ? rows : cols;
  const long long total = (long long)out_rows * ld_dst;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / ld_dst), c = (int)(t - (long long)r * ld_dst);
    float x = 0.f;
    if (c < out_cols) x = transpose ? src[(long long)c * ld_src + r] : src[(long long)r * ld_src + c];
    const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    hi[t] = h;
    lo[t] = x - h;
  }
}

// gZ = gout * act'(Z) from the saved OUTPUT (sign(out) == sign(Z) for leaky / relu / identity),
// written pre-split and padded: the A operand of both backward GEMMs.
__global__ void __launch_bounds__(256) act_bwd_split_kernel(float *__restrict__ gz_hi, float *__restrict__ gz_lo, int ldz,
                                                            const float *__restrict__ gout, const float *__restrict__ out,
                                                            int M, int U, float slope) {
  const long long total = (long long)M * ldz;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / ldz), c = (int)(t - (long long)r * ldz);
    float x = 0.f;
    if (c < U) {
      const long long o = (long long)r * U + c;
      const float gval = __ldg(gout + o);
      x = __ldg(out + o) > 0.f ? gval : slope * gval;
    }
    if (gz_lo) {
      const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
      gz_hi[t] = h;
      gz_lo[t] = x - h;
    } else {
      gz_hi[t] = x;   // plain fp32 for the in-kernel-split GEMM
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Encoded tensor maps are cached per thread, keyed by (pointer, shape, box): a training loop calls the same
// GEMMs on the same caching-allocator blocks every step, and cuTensorMapEncodeTiled costs ~1 us each.
struct MapKey {
  const void *ptr; int d0, d1, ld, box, kind;
  bool operator==(const MapKey &o) const { return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && box == o.box && kind == o.kind; }
};
struct MapSlot { MapKey key; CUtensorMap map; bool used; };
constexpr int kMapCacheSlots = 256;
static thread_local MapSlot g_map_cache[kMapCacheSlots];
static MapSlot *map_slot(const MapKey &k) {
  uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
  h ^= ((uint64_t)(uint32_t)k.d0 << 32 | (uint32_t)k.d1) * 0xC2B2AE3D27D4EB4Full;
  h ^= ((uint64_t)(uint32_t)k.ld << 32 | (uint32_t)(k.box * 4 + k.kind)) * 0x165667B19E3779F9ull;
  return &g_map_cache[(h >> 40) % kMapCacheSlots];
}

// K-major operand X[rows, K] (row-major, ld floats): 2-D map, box = 32 floats x box_rows
static int make_map_kmajor(CUtensorMap *m, const float *x, int rows, int K, int ld, int box_rows) {
  const MapKey key{x, rows, K, ld, box_rows, 0};
  MapSlot *slot = map_slot(key);
  if (slot->used && slot->key == key) { *m = slot->map; return SG_OK; }
  const int rc_ = [&]() -> int {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(SG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SG_ERR_CUDA, "cuTensorMapEncodeTiled (K-major) failed with CUresult %d", (int)r);
  return SG_OK;
  }();
  if (rc_ == SG_OK) { slot->key = key; slot->map = *m; slot->used = true; }
  return rc_;
}

// MN-major operand X[K, mn] (row-major, ld floats): 2-D map (mn, K), box = 32 floats x 32 k-rows = one
// SWIZZLE_128B_ATOM_32B box; the kernel places the atoms of a tile 4096 B apart -> smem [atom][k-row][32 floats]
static int make_map_mnmajor(CUtensorMap *m, const float *x, int K, int mn, int ld) {
  const MapKey key{x, K, mn, ld, 32, 1};
  MapSlot *slot = map_slot(key);
  if (slot->used && slot->key == key) { *m = slot->map; return SG_OK; }
  const int rc_ = [&]() -> int {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(SG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)mn, (cuuint64_t)K};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)kBK};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SG_ERR_CUDA, "cuTensorMapEncodeTiled (MN-major) failed with CUresult %d", (int)r);
  return SG_OK;
  }();
  if (rc_ == SG_OK) { slot->key = key; slot->map = *m; slot->used = true; }
  return rc_;
}

template <int STAGES, bool MN>
static int launch_gemm_pair(const CUtensorMap (&maps)[4], GemmArgs g, int splits, cudaStream_t st) {
  constexpr int smem = STAGES * 2 * (kBM * kBK * 4 + 128 * kBK * 4) + kEpiWarps * 32 * 33 * 4 + 1024;
  SG_CUDA(cudaFuncSetAttribute(tf32x3_gemm_pair_kernel<STAGES, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  g.tiles_m = ceil_div(g.M, 256);
  g.tiles_n = ceil_div(g.N, 256);
  g.splits = splits;
  const long long n_tiles = (long long)g.tiles_m * g.tiles_n * splits;
  const int max_pairs = num_sms() / 2;
  const int pairs = (int)(n_tiles < max_pairs ? n_tiles : max_pairs);
  tf32x3_gemm_pair_kernel<STAGES, MN><<<2 * pairs, kGemmThreads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], g);
  SG_LAUNCHED("tf32x3_gemm_pair_kernel");
  return SG_OK;
}

template <int STAGES, bool MN, bool SPLIT_B>
static int launch_gemm_split(const CUtensorMap (&maps)[4], GemmArgs g, int splits, cudaStream_t st) {
  constexpr int smem = STAGES * 2 * (kBM * kBK * 4 + 128 * kBK * 4) + kEpiWarps * 32 * 33 * 4 + 1024;
  SG_CUDA(cudaFuncSetAttribute(tf32x3_gemm_split_kernel<STAGES, MN, SPLIT_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  g.tiles_m = ceil_div(g.M, 256);
  g.tiles_n = ceil_div(g.N, 256);
  g.splits = splits;
  const long long n_tiles = (long long)g.tiles_m * g.tiles_n * splits;
  const int max_pairs = num_sms() / 2;
  const int pairs = (int)(n_tiles < max_pairs ? n_tiles : max_pairs);
  tf32x3_gemm_split_kernel<STAGES, MN, SPLIT_B><<<2 * pairs, kSplitThreads, smem, st>>>(maps[0], maps[2], maps[3], g);
  SG_LAUNCHED("tf32x3_gemm_split_kernel");
  return SG_OK;
}

static inline int grid_ew(long long n) {
  long long gsz = ceil_div<long long>(n > 0 ? n : 1, 256);
  long long cap = (long long)num_sms() * 32;
  return (int)(gsz < cap ? gsz : cap);
}

}  // namespace sg

using namespace sg;

extern "C" {

size_t sg_gemm_split_ws_bytes(int M, int N, int splits) {
  if (M <= 0 || N <= 0 || splits <= 1) return 0;
  return (size_t)splits * (size_t)M * (size_t)N * sizeof(float);
}

int sg_gemm_tf32x3(float *D, int ldd, const float *A_hi, const float *A_lo, int lda, const float *B_hi,
                   const float *B_lo, int ldb, int M, int N, int K, int mn_major, int epilogue, float slope,
                   const float *bias, int splits, float *split_ws, sg_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_REQUIRE(M > 0 && N > 0 && K > 0, "sg_gemm_tf32x3: bad sizes M=%d N=%d K=%d", M, N, K);
  SG_REQUIRE(D && A_hi && B_hi, "sg_gemm_tf32x3: null pointer");
  SG_REQUIRE(B_lo || !A_lo, "sg_gemm_tf32x3: a raw B operand (B_lo == NULL) needs a raw A operand (A_lo == NULL) as well");
  SG_REQUIRE(epilogue == 0 || epilogue == 1, "sg_gemm_tf32x3: bad epilogue %d", epilogue);
  SG_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0, "sg_gemm_tf32x3: operand leading dimensions must be multiples of 4 floats (TMA strides)");
  const uintptr_t al = reinterpret_cast<uintptr_t>(A_hi) | reinterpret_cast<uintptr_t>(A_lo) |
                       reinterpret_cast<uintptr_t>(B_hi) | reinterpret_cast<uintptr_t>(B_lo);
  SG_REQUIRE((al & 15) == 0, "sg_gemm_tf32x3: operands must be 16-byte aligned");
  if (splits < 1) splits = 1;
  SG_REQUIRE(splits == 1 || (split_ws && epilogue == 0 && !bias), "sg_gemm_tf32x3: split-K needs a workspace and the plain epilogue");
  GemmArgs g;
  g.M = M; g.N = N; g.epi = epilogue; g.slope = slope; g.bias = bias;
  g.chain_kb = dev_option(SG_DEV_GEMM_CHAIN) > 0 ? dev_option(SG_DEV_GEMM_CHAIN) : kChainKBlocks;
  g.relaxed_arrive = dev_option(SG_DEV_GEMM_ARRIVE) == 0;
  g.trace = dev_option(SG_DEV_GEMM_TRACE);
  g.kb_total = ceil_div(K, kBK);
  g.kb_per_split = ceil_div(g.kb_total, splits);
  splits = ceil_div(g.kb_total, g.kb_per_split);  // no empty split
  if (splits > 1) { g.D = split_ws; g.ldd = N; g.split_stride = (long long)M * N; }
  else { g.D = D; g.ldd = ldd; g.split_stride = 0; }
  CUtensorMap maps[4];
  int rc;
  if (mn_major) {
    SG_REQUIRE(M <= lda && N <= ldb, "sg_gemm_tf32x3: MN-major leading dimensions too small");
    if ((rc = make_map_mnmajor(&maps[0], A_hi, K, M, lda)) != SG_OK) return rc;
    if ((rc = make_map_mnmajor(&maps[1], A_lo ? A_lo : A_hi, K, M, lda)) != SG_OK) return rc;
    if ((rc = make_map_mnmajor(&maps[2], B_hi, K, N, ldb)) != SG_OK) return rc;
    if ((rc = make_map_mnmajor(&maps[3], B_lo ? B_lo : B_hi, K, N, ldb)) != SG_OK) return rc;
    if (!A_lo) rc = B_lo ? launch_gemm_split<3, true, false>(maps, g, splits, st) : launch_gemm_split<3, true, true>(maps, g, splits, st);
    else rc = launch_gemm_pair<3, true>(maps, g, splits, st);
    if (rc != SG_OK) return rc;
  } else {
    if ((rc = make_map_kmajor(&maps[0], A_hi, M, K, lda, kBM)) != SG_OK) return rc;
    if ((rc = make_map_kmajor(&maps[1], A_lo ? A_lo : A_hi, M, K, lda, kBM)) != SG_OK) return rc;
    if ((rc = make_map_kmajor(&maps[2], B_hi, N, K, ldb, 128)) != SG_OK) return rc;   // a CTA of a pair loads half of the B tile
    if ((rc = make_map_kmajor(&maps[3], B_lo ? B_lo : B_hi, N, K, ldb, 128)) != SG_OK) return rc;
    if (!A_lo) rc = B_lo ? launch_gemm_split<3, false, false>(maps, g, splits, st) : launch_gemm_split<3, false, true>(maps, g, splits, st);
    else rc = launch_gemm_pair<3, false>(maps, g, splits, st);
    if (rc != SG_OK) return rc;
  }
  if (splits > 1) {
    splitk_reduce_kernel<<<grid_ew((long long)M * N), 256, 0, st>>>(D, ldd, split_ws, M, N, splits);
    SG_LAUNCHED("splitk_reduce_kernel");
  }
  return SG_OK;
}

/* Development: copy the 8 trace counters (see g_gemm_trace) to the host and clear them.  Synchronises the device. */
int sg_gemm_trace_read(unsigned long long *host8) {
  SG_REQUIRE(host8, "sg_gemm_trace_read: null pointer");
  SG_CUDA(cudaDeviceSynchronize());
  SG_CUDA(cudaMemcpyFromSymbol(host8, g_gemm_trace, 8 * sizeof(unsigned long long)));
  unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  SG_CUDA(cudaMemcpyToSymbol(g_gemm_trace, zero, sizeof(zero)));
  return SG_OK;
}

/* Measurement aid (tools/gemm_bench.py): TMA delivery rate per SM, see tma_probe_kernel. */
int sg_tma_probe(const float *src, int rows, int K, int ld, int stages, int boxes, int iters, int blocks, sg_stream_t stream) {
  // encoding kept in one int to leave the signature alone: stages = real stages + 100 * (producer warps - 1)
  const int producers = stages / 100 + 1;
  stages %= 100;
  SG_REQUIRE(producers >= 1 && producers <= 4 && producers <= boxes, "sg_tma_probe: 1..4 producer warps, at most one per box");
  const int same_boxes = iters < 0;       // negative iters: all CTAs read the same sequence of boxes (hot lines)
  if (iters < 0) iters = -iters;
  SG_REQUIRE(src && rows >= kBM && K >= kBK && (ld & 3) == 0 && stages >= 1 && stages <= 16 && boxes >= 1 && iters >= 1 && blocks >= 1,
             "sg_tma_probe: bad arguments");
  const int smem = stages * boxes * kBM * kBK * 4 + 1024;
  SG_REQUIRE(smem <= 227 * 1024, "sg_tma_probe: ring of %d bytes does not fit shared memory", smem);
  CUtensorMap map;
  int rc = make_map_kmajor(&map, src, rows, K, ld, kBM);
  if (rc != SG_OK) return rc;
  SG_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tma_probe_kernel<<<blocks, 160, smem, (cudaStream_t)stream>>>(map, rows, K / kBK, stages, boxes, iters, same_boxes, producers);
  SG_LAUNCHED("tma_probe_kernel");
  return SG_OK;
}

int sg_split_tf32(float *hi, float *lo, int ld_dst, const float *src, int rows, int cols, int ld_src, int transpose,
                  sg_stream_t stream) {
  SG_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols, "sg_split_tf32: bad sizes");
  SG_REQUIRE(ld_dst >= (transpose ? rows : cols), "sg_split_tf32: ld_dst too small");
  if (rows == 0 || cols == 0) return SG_OK;
  SG_REQUIRE(hi && lo && src, "sg_split_tf32: null pointer");
  const long long total = (long long)(transpose ? cols : rows) * ld_dst;
  split_tf32_kernel<<<grid_ew(total), 256, 0, (cudaStream_t)stream>>>(hi, lo, ld_dst, src, rows, cols, ld_src, transpose);
  SG_LAUNCHED("split_tf32_kernel");
  return SG_OK;
}

int sg_act_bwd_split(float *gz_hi, float *gz_lo, int ldz, const float *gout, const float *out, int M, int U,
                     float slope, sg_stream_t stream) {
  SG_REQUIRE(M >= 0 && U >= 0 && ldz >= U, "sg_act_bwd_split: bad sizes");
  if (M == 0 || U == 0) return SG_OK;
  SG_REQUIRE(gz_hi && gout && out, "sg_act_bwd_split: null pointer");
  act_bwd_split_kernel<<<grid_ew((long long)M * ldz), 256, 0, (cudaStream_t)stream>>>(gz_hi, gz_lo, ldz, gout, out, M, U, slope);
  SG_LAUNCHED("act_bwd_split_kernel");
  return SG_OK;
}

}  // extern "C"
