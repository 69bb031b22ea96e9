// The hot kernel: CSR weighted gather-accumulate
//     out[seg, :] (=|+=) sum_{p in seg} w[p] * src[idx[p], :]
// used for seg_weighted_pool forward (A2), its data-gradient over the transposed pattern (A3),
// seg_pool sum/mean (A5) and the fused multi-relation aggregation forward/backward (A1).
//
// Mapping (HBM/L2-bandwidth bound, ~0.5 flop per gathered byte, so no tensor cores here):
//   * a GROUP of LPR lanes owns one work item (a run of <= chunk edges of one segment) and
//     covers one source row with 16-byte loads: for D=64 a row is 256 B = 16 lanes x float4,
//     so a warp keeps two independent items in flight and every row load is one fully
//     used 128-B line pair
//   * UNROLL independent row loads are issued before any FMA (memory-level parallelism)
//   * heavy-tailed degree distributions are balanced by the plan (plan.cu): long segments are
//     cut into chunks whose partial rows are combined by a second, tiny kernel in a fixed
//     order -> no float atomics, bit-identical results run to run (the reference also avoids
//     atomics: seg_op.cu:747-790)
// Reference kernels replaced: SegTakeKCorrBackwardEmbed1Kernel (seg_op.cu:682-722),
// SegTakeKCorrBackwardEmbed2Kernel (seg_op.cu:747-790), SegPoolKernel sum/mean
// (seg_op.cu:1057-1135), SegPoolBackwardKernel sum/mean (seg_op.cu:1171-1215).
#include "common.cuh"
#include "gather.cuh"

namespace sg {

template <int VEC> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<1> { using T = float; };

template <int VEC>
__device__ __forceinline__ void ld_row(float (&v)[VEC], const float *p) {
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2 *>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}

template <int VEC>
__device__ __forceinline__ void ld_plain(float (&v)[VEC], const float *p) {
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = *reinterpret_cast<const float2 *>(p);
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = *p;
  }
}

template <int VEC>
__device__ __forceinline__ void st_row(float *p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else if constexpr (VEC == 2) *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
  else *p = v[0];
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// r = seg / n_out_rows for 0 <= seg < 2^31 by multiply-high (Granlund-Montgomery round-up method:
// shift = ceil(log2 d), magic = floor(2^32 (2^shift - d) / d) + 1, q = (mulhi(magic, n) + n) >> shift; the
// sum cannot overflow because mulhi(magic, n) < n < 2^31).  The hardware has no integer divide: the
// compiler's expansion costs ~30 instructions per segment, which the short user-side segments (11 edges)
// paid twice — row offset and weight-sum offset.
__device__ __forceinline__ int seg_rel(const GatherArgs &a, int seg) {
  return (int)((__umulhi(a.div_magic, (uint32_t)seg) + (uint32_t)seg) >> a.div_shift);
}

__device__ __forceinline__ void store_wsum_at(const GatherArgs &a, int i, int r, float wacc) {
  const long long o = (long long)i * a.wsum_ld + r;
  if (a.wsum_lo) {
    const float h = tf32_hi(wacc);
    a.wsum[o] = h;
    a.wsum_lo[o] = wacc - h;
  } else {
    a.wsum[o] = wacc;
  }
}

__device__ __forceinline__ void store_wsum(const GatherArgs &a, int seg, float wacc) {
  const int r = seg_rel(a, seg);
  store_wsum_at(a, seg - r * a.n_out_rows, r, wacc);
}

__device__ __forceinline__ long long out_offset(const GatherArgs &a, int seg) {
  if (a.n_out_rows == a.n_seg) return (long long)seg * a.ld_out;
  const int r = seg_rel(a, seg), i = seg - r * a.n_out_rows;
  return (long long)i * a.ld_out + (long long)r * a.F;
}

// Destination of output row `row` when the rows are scattered to their owners' staging buffers (peer memory)
__device__ __forceinline__ float *peer_row(const GatherArgs &a, int row) {
  float *base = a.peer_out[0];
  int lo = 0;
#pragma unroll
  for (int q = 1; q < SG_MAX_PEERS + 1; ++q)
    if (q < a.peer_world && row >= a.peer_lo[q]) { base = a.peer_out[q]; lo = a.peer_lo[q]; }
  return base + (long long)(row - lo) * a.ld_out;
}

template <int VEC, int LPR, int NV, int UNROLL>
__global__ void __launch_bounds__(256) gather_rows_kernel(const GatherArgs a) {
  const int lane = threadIdx.x & (LPR - 1);
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int n_groups = (gridDim.x * blockDim.x) / LPR;
  const int k = blockIdx.y;
  const int col0 = blockIdx.z * (LPR * NV * VEC);  // column chunk for very wide rows

  const float *__restrict__ src = a.src + (long long)k * a.src_batch_stride;
  const float *__restrict__ w = a.w ? a.w + (long long)k * a.w_batch_stride : nullptr;
  float *__restrict__ out = a.out + (long long)k * a.out_batch_stride;
  float *__restrict__ partial = a.partial ? a.partial + (long long)k * a.partial_batch_stride : nullptr;
  const int32_t *__restrict__ idx = a.idx;
  const int32_t *__restrict__ perm = a.perm;

  const int n_items = a.hdr ? a.hdr->n_items : a.n_seg;

  for (int it = group; it < n_items; it += n_groups) {
    int4 d;
    if (a.hdr) {
      d = __ldg(a.items + it);
    } else {
      d = make_int4(__ldg(a.indptr + it), __ldg(a.indptr + it + 1), it, -1);
    }
    float acc[NV][VEC];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[v][e] = 0.f;
    float wacc = 0.f;

    for (int p = d.x; p < d.y; p += UNROLL) {
      int id[UNROLL];
      float wv[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const bool ok = p + u < d.y;
        id[u] = ok ? __ldg(idx + p + u) : -1;
        wv[u] = 1.f;
        if (w) wv[u] = ok ? (perm ? __ldg(w + __ldg(perm + p + u)) : __ldg(w + p + u)) : 0.f;
        if (a.inv_len_indptr && ok)
          wv[u] = 1.f / (float)(__ldg(a.inv_len_indptr + id[u] + 1) - __ldg(a.inv_len_indptr + id[u]));
      }
      float val[UNROLL][NV][VEC];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = col0 + (v * LPR + lane) * VEC;
          if (id[u] >= 0 && c < a.F) {
            ld_row<VEC>(val[u][v], src + (long long)id[u] * a.ld_src + c);
          } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) val[u][v][e] = 0.f;
          }
        }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (id[u] >= 0) {
          wacc += wv[u];
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[v][e] = fmaf(wv[u], val[u][v][e], acc[v][e]);
        }
      }
    }

    if (d.w >= 0) {  // piece of a split segment: park the partial row
      float *prow = partial + (long long)d.w * a.F;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = col0 + (v * LPR + lane) * VEC;
        if (c < a.F) st_row<VEC>(prow + c, acc[v]);
      }
      if (a.partial_wsum && lane == 0 && blockIdx.z == 0) a.partial_wsum[d.w] = wacc;
    } else {
      float *orow = out + out_offset(a, d.z);
      const float inv = a.mean && d.y > d.x ? 1.f / (float)(d.y - d.x) : 1.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = col0 + (v * LPR + lane) * VEC;
        if (c < a.F) {
          float r[VEC];
#pragma unroll
          for (int e = 0; e < VEC; ++e) r[e] = a.mean ? acc[v][e] * inv : acc[v][e];
          if (a.req == SG_REQ_ADD) {
            float o[VEC];
            ld_plain<VEC>(o, orow + c);
#pragma unroll
            for (int e = 0; e < VEC; ++e) r[e] += o[e];
          }
          st_row<VEC>(orow + c, r);
        }
      }
      if (a.wsum && lane == 0 && blockIdx.z == 0) store_wsum(a, d.z, wacc);
    }
  }
}


// ------------------------------------------------------------------------------------------
// Fast path: rows of exactly LPR*4 floats (F in {16,32,64,128}), every per-call option a
// template parameter so the edge loop is  idx/w loads -> address -> LDG.128 -> 4 FFMA  and
// nothing else (the generic kernel above spends ~50 issued instructions per edge on runtime
// option checks; this one ~8).  Full UNROLL-edge batches run unpredicated; one predicated
// batch handles the tail.
//   WMODE 0: weight 1   1: w[p]   2: w[perm[p]]   3: 1/len(segment idx[p]) (avg-pool backward)
// ------------------------------------------------------------------------------------------
template <int WMODE>
__device__ __forceinline__ float edge_weight(const GatherArgs &a, const float *__restrict__ w, int p, int id) {
  if constexpr (WMODE == 0) return 1.f;
  else if constexpr (WMODE == 1) return __ldg(w + p);
  else if constexpr (WMODE == 2) return __ldg(w + __ldg(a.perm + p));
  else return 1.f / (float)(__ldg(a.inv_len_indptr + id + 1) - __ldg(a.inv_len_indptr + id));
}

// LPR lanes x NV float4 per lane cover one row (F = LPR * NV * 4): lane l loads the float4 at column
// (v * LPR + l) * 4, so each of the NV load instructions of a group touches LPR * 16 contiguous bytes.
// NV = 2 at D = 64 puts FOUR work items in a warp instead of two: half the issued instructions per
// edge, which is what bounds the short-segment (user-side) launch.
// __launch_bounds__(256, 1): no register cap.  The compiler then takes 46 (batches of 4) / 64 (batches of 8)
// registers and issues the loads of a batch further ahead; with the 40-register build (6 resident blocks
// instead of 4) the long-segment launches were 5 % slower, and a 32-register cap (64 resident warps) did not
// help the short-segment launch either — none of them is occupancy-bound (profiles/r01_summary.md §H).
template <int LPR, int NV, int UNROLL, int WMODE, bool WSUM, bool PLAIN, bool PEER>
__global__ void __launch_bounds__(256, 1) gather_rows_fast_kernel(const GatherArgs a) {
  constexpr int F = LPR * NV * 4;
  const int lane = threadIdx.x & (LPR - 1);
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int n_groups = (gridDim.x * blockDim.x) / LPR;
  const int k = blockIdx.y;

  const float4 *__restrict__ src = reinterpret_cast<const float4 *>(a.src + (long long)k * a.src_batch_stride) + lane;
  const float *__restrict__ w = WMODE == 1 || WMODE == 2 ? a.w + (long long)k * a.w_batch_stride : nullptr;
  float *__restrict__ out = a.out + (long long)k * a.out_batch_stride;
  const int32_t *__restrict__ idx = a.idx;
  constexpr int ld4 = F / 4;
  const int n_items = a.hdr ? a.hdr->n_items : a.n_seg;

  for (int it = group; it < n_items; it += n_groups) {
    int4 d;
    if (a.hdr) d = __ldg(a.items + it);
    else d = make_int4(__ldg(a.indptr + it), __ldg(a.indptr + it + 1), it, -1);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    float wacc = 0.f;
    int p = d.x;
    for (; p + UNROLL <= d.y; p += UNROLL) {
      int id[UNROLL];
      float wv[UNROLL];
      float4 val[UNROLL][NV];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) id[u] = __ldg(idx + p + u);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int v = 0; v < NV; ++v) val[u][v] = __ldg(src + (long long)id[u] * ld4 + v * LPR);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) wv[u] = edge_weight<WMODE>(a, w, p + u, id[u]);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if constexpr (WSUM) wacc += wv[u];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          acc[v].x = fmaf(wv[u], val[u][v].x, acc[v].x);
          acc[v].y = fmaf(wv[u], val[u][v].y, acc[v].y);
          acc[v].z = fmaf(wv[u], val[u][v].z, acc[v].z);
          acc[v].w = fmaf(wv[u], val[u][v].w, acc[v].w);
        }
      }
    }
    if (p < d.y) {  // tail: one predicated batch
      int id[UNROLL];
      float wv[UNROLL];
      float4 val[UNROLL][NV];
#pragma unroll
      for (int u = 0; u < UNROLL - 1; ++u) id[u] = p + u < d.y ? __ldg(idx + p + u) : -1;
#pragma unroll
      for (int u = 0; u < UNROLL - 1; ++u)
#pragma unroll
        for (int v = 0; v < NV; ++v)
          val[u][v] = id[u] >= 0 ? __ldg(src + (long long)id[u] * ld4 + v * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < UNROLL - 1; ++u) wv[u] = id[u] >= 0 ? edge_weight<WMODE>(a, w, p + u, id[u]) : 0.f;
#pragma unroll
      for (int u = 0; u < UNROLL - 1; ++u) {
        if (id[u] >= 0) {  // keeps the addition order of the serial loop and never touches a masked row
          if constexpr (WSUM) wacc += wv[u];
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            acc[v].x = fmaf(wv[u], val[u][v].x, acc[v].x);
            acc[v].y = fmaf(wv[u], val[u][v].y, acc[v].y);
            acc[v].z = fmaf(wv[u], val[u][v].z, acc[v].z);
            acc[v].w = fmaf(wv[u], val[u][v].w, acc[v].w);
          }
        }
      }
    }

    if (d.w >= 0) {  // piece of a split segment: park the partial row
      float *prow = a.partial + (long long)k * a.partial_batch_stride + (long long)d.w * F;
#pragma unroll
      for (int v = 0; v < NV; ++v) reinterpret_cast<float4 *>(prow)[v * LPR + lane] = acc[v];
      if (WSUM && lane == 0) a.partial_wsum[d.w] = wacc;
    } else {
      int rel = 0, row = d.z;  // segment = rel * n_out_rows + row (relation-major concatenated CSRs)
      if (a.n_out_rows != a.n_seg) { rel = seg_rel(a, d.z); row = d.z - rel * a.n_out_rows; }
      float *orow;
      if constexpr (PEER) orow = peer_row(a, row);   // PEER implies PLAIN and n_out_rows == n_seg
      else orow = out + ((long long)row * a.ld_out + rel * F);
      float inv = 1.f;
      if constexpr (!PLAIN) inv = (a.mean && d.y > d.x) ? 1.f / (float)(d.y - d.x) : 1.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float4 r = acc[v];
        if constexpr (!PLAIN) {  // PLAIN: req == write and no mean (the fused aggregation's two launches)
          if (a.mean) { r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv; }
          if (a.req == SG_REQ_ADD) {
            const float4 o = reinterpret_cast<const float4 *>(orow)[v * LPR + lane];
            r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
          }
        }
        if (!PEER && a.out_lo) {
          const float4 hi = make_float4(tf32_hi(r.x), tf32_hi(r.y), tf32_hi(r.z), tf32_hi(r.w));
          reinterpret_cast<float4 *>(orow)[v * LPR + lane] = hi;
          reinterpret_cast<float4 *>(a.out_lo + (orow - a.out))[v * LPR + lane] =
              make_float4(r.x - hi.x, r.y - hi.y, r.z - hi.z, r.w - hi.w);
        } else {
          reinterpret_cast<float4 *>(orow)[v * LPR + lane] = r;
        }
      }
      if (WSUM && lane == 0) store_wsum_at(a, row, rel, wacc);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Staged variant (development option SG_DEV_GATHER_VARIANT = 1; measured against the default in
// profiles/r02_summary.md): ONE WARP per work item; lane 0 bulk-copies the item's index and weight
// segments global -> shared with cp.async.bulk (the TMA unit's 1-D copy, completion on an mbarrier) one
// item ahead of the row loads, the two half-warps then walk alternate 8-edge batches reading indices /
// weights from shared memory (broadcast reads) and their two partial rows are combined with one
// shuffle-xor before the store.  Rows of 64 floats, weights w[p], write semantics (the fused
// aggregation's launches); only used for long work items (>= 16 edges on average).
// ------------------------------------------------------------------------------------------
constexpr int kStageElems = 264;   // 256-edge chunk + alignment slack on both ends (16-byte bulk copies)

__device__ __forceinline__ uint32_t gsmem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool WSUM>
__global__ void __launch_bounds__(256) gather_rows_staged_kernel(const GatherArgs a) {
  __shared__ __align__(16) int32_t s_idx[8][2][kStageElems];
  __shared__ __align__(16) float s_w[8][2][kStageElems];
  __shared__ uint64_t s_bar[8][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  const int gwarp = blockIdx.x * 8 + warp, n_warps = gridDim.x * 8;
  const int n_items = a.hdr->n_items;
  const int nnz4 = (a.hdr->nnz + 3) & ~3;
  const float4 *__restrict__ src = reinterpret_cast<const float4 *>(a.src) + l16;

  if (lane == 0) {
    for (int b = 0; b < 2; ++b)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gsmem_u32(&s_bar[warp][b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  auto prefetch = [&](int it, int buf) {   // lane 0 only
    const int4 d = __ldg(a.items + it);
    const int p0 = d.x & ~3;
    int p1 = (d.y + 3) & ~3;
    if (p1 > nnz4) p1 = nnz4;
    const uint32_t bytes = (uint32_t)(p1 - p0) * 4u;
    const uint32_t bar = gsmem_u32(&s_bar[warp][buf]);
    if (bytes == 0) {   // empty item: complete the phase without a copy
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
      return;
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(gsmem_u32(&s_idx[warp][buf][0])), "l"(a.idx + p0), "r"(bytes), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(gsmem_u32(&s_w[warp][buf][0])), "l"(a.w + p0), "r"(bytes), "r"(bar) : "memory");
  };

  int k = 0;
  if (gwarp < n_items && lane == 0) prefetch(gwarp, 0);
  for (int it = gwarp; it < n_items; it += n_warps, ++k) {
    const int buf = k & 1;
    const uint32_t phase = (uint32_t)(k >> 1) & 1u;
    const int4 d = __ldg(a.items + it);
    if (lane == 0 && it + n_warps < n_items) prefetch(it + n_warps, buf ^ 1);
    {
      const uint32_t bar = gsmem_u32(&s_bar[warp][buf]);
      asm volatile(
          "{\n\t"
          ".reg .pred P1;\n\t"
          "LAB_WAIT:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
          "@P1 bra DONE;\n\t"
          "bra LAB_WAIT;\n\t"
          "DONE:\n\t"
          "}" ::"r"(bar), "r"(phase) : "memory");
    }
    const int32_t *sidx = &s_idx[warp][buf][0] + (d.x & 3);
    const float *sw = &s_w[warp][buf][0] + (d.x & 3);
    const int n = d.y - d.x;
    float4 acc[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    float wacc = 0.f;
    // batches of 16 edges: half-warp h takes edges [e0 + 8h, e0 + 8h + 8)
    for (int e0 = 0; e0 < n; e0 += 16) {
      const int eb = e0 + half * 8;
      int id[8];
      float wv[8];
      float4 val[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) id[u] = eb + u < n ? sidx[eb + u] : -1;
#pragma unroll
      for (int u = 0; u < 8; ++u) val[u] = id[u] >= 0 ? __ldg(src + (long long)id[u] * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) wv[u] = id[u] >= 0 ? sw[eb + u] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if constexpr (WSUM) wacc += wv[u];
        acc[u & 3].x = fmaf(wv[u], val[u].x, acc[u & 3].x);
        acc[u & 3].y = fmaf(wv[u], val[u].y, acc[u & 3].y);
        acc[u & 3].z = fmaf(wv[u], val[u].z, acc[u & 3].z);
        acc[u & 3].w = fmaf(wv[u], val[u].w, acc[u & 3].w);
      }
    }
    float4 r;
    r.x = (acc[0].x + acc[1].x) + (acc[2].x + acc[3].x);
    r.y = (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y);
    r.z = (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z);
    r.w = (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w);
    r.x += __shfl_xor_sync(0xffffffffu, r.x, 16);
    r.y += __shfl_xor_sync(0xffffffffu, r.y, 16);
    r.z += __shfl_xor_sync(0xffffffffu, r.z, 16);
    r.w += __shfl_xor_sync(0xffffffffu, r.w, 16);
    if constexpr (WSUM) wacc += __shfl_xor_sync(0xffffffffu, wacc, 16);
    if (half == 0) {
      if (d.w >= 0) {
        reinterpret_cast<float4 *>(a.partial + (long long)d.w * 64)[l16] = r;
        if (WSUM && l16 == 0) a.partial_wsum[d.w] = wacc;
      } else {
        int rel = 0, row = d.z;
        if (a.n_out_rows != a.n_seg) { rel = seg_rel(a, d.z); row = d.z - rel * a.n_out_rows; }
        float *orow = a.out + ((long long)row * a.ld_out + rel * 64);
        if (a.out_lo) {
          const float4 hi = make_float4(tf32_hi(r.x), tf32_hi(r.y), tf32_hi(r.z), tf32_hi(r.w));
          reinterpret_cast<float4 *>(orow)[l16] = hi;
          reinterpret_cast<float4 *>(a.out_lo + (orow - a.out))[l16] = make_float4(r.x - hi.x, r.y - hi.y, r.z - hi.z, r.w - hi.w);
        } else {
          reinterpret_cast<float4 *>(orow)[l16] = r;
        }
        if (WSUM && l16 == 0) store_wsum_at(a, row, rel, wacc);
      }
    }
    __syncwarp();   // every lane is done with this buffer before lane 0 refills it (two items ahead)
  }
}

// Second pass: fixed-order sum of the partial rows of every split segment.
template <int VEC, int LPR, int NV, bool PEER = false>
__global__ void __launch_bounds__(256) combine_partials_kernel(const GatherArgs a) {
  const int lane = threadIdx.x & (LPR - 1);
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int n_groups = (gridDim.x * blockDim.x) / LPR;
  const int k = blockIdx.y;
  const int col0 = blockIdx.z * (LPR * NV * VEC);
  float *__restrict__ out = a.out + (long long)k * a.out_batch_stride;
  const float *__restrict__ partial = a.partial + (long long)k * a.partial_batch_stride;
  const int n_long = a.hdr->n_long;
  for (int li = group; li < n_long; li += n_groups) {
    const int4 d = __ldg(a.longs + li);  // {segment, first slot, count}
    float acc[NV][VEC];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[v][e] = 0.f;
    float wacc = 0.f;
    // CU partial rows are requested before any is added (the hottest item of a rating graph is cut into
    // hundreds of partials: one dependent load per partial made this pass 35 us); additions stay in slot order
    constexpr int CU = NV * VEC <= 4 ? 8 : 2;
    for (int s0 = 0; s0 < d.z; s0 += CU) {
      float t[CU][NV][VEC];
      float tw[CU];
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const bool ok = s0 + u < d.z;
        const float *prow = partial + (long long)(d.y + s0 + u) * a.F;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = col0 + (v * LPR + lane) * VEC;
          if (ok && c < a.F) {
            ld_plain<VEC>(t[u][v], prow + c);
          } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) t[u][v][e] = 0.f;
          }
        }
        tw[u] = (ok && a.partial_wsum) ? a.partial_wsum[d.y + s0 + u] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        if (s0 + u < d.z) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[v][e] += t[u][v][e];
          wacc += tw[u];
        }
      }
    }
    float *orow;
    if constexpr (PEER) orow = peer_row(a, d.x);
    else orow = out + out_offset(a, d.x);
    float inv = 1.f;
    if (a.mean) {
      const int len = __ldg(a.indptr + d.x + 1) - __ldg(a.indptr + d.x);
      inv = 1.f / (float)len;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = col0 + (v * LPR + lane) * VEC;
      if (c < a.F) {
        float r[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) r[e] = a.mean ? acc[v][e] * inv : acc[v][e];
        if (a.req == SG_REQ_ADD) {
          float o[VEC];
          ld_plain<VEC>(o, orow + c);
#pragma unroll
          for (int e = 0; e < VEC; ++e) r[e] += o[e];
        }
        if (!PEER && a.out_lo) {
          float h[VEC], l[VEC];
#pragma unroll
          for (int e = 0; e < VEC; ++e) { h[e] = tf32_hi(r[e]); l[e] = r[e] - h[e]; }
          st_row<VEC>(orow + c, h);
          st_row<VEC>(a.out_lo + (orow - a.out) + c, l);
        } else {
          st_row<VEC>(orow + c, r);
        }
      }
    }
    if (a.wsum && lane == 0 && blockIdx.z == 0) store_wsum(a, d.x, wacc);
  }
}

// Grid-stride launches use at most 32 blocks per SM (5-6 resident at a time): work items differ in length
// (chunks of up to 256 edges), so several waves of smaller blocks balance better than one resident wave —
// measured on the ML-10M shape, sum of the four launches: 0.99 ms at 6 blocks/SM, 0.86 at 16, 0.82 at
// 24..64, 0.86 with one item per group.  (Development option SG_DEV_GATHER_GRID overrides it for sweeps.)
static long long grid_cap() {
  const int v = dev_option(SG_DEV_GATHER_GRID);
  return (long long)num_sms() * (v > 0 ? v : 32);
}

template <int VEC, int LPR, int NV>
static int launch_gather(const GatherArgs &a, int K, int n_items_cap, int n_long_cap, cudaStream_t st) {
  constexpr int UNROLL = NV >= 2 ? 2 : 4;
  constexpr int kThreads = 256;
  constexpr int groups_per_block = kThreads / LPR;
  const int col_chunks = ceil_div(a.F, LPR * NV * VEC);
  long long blocks = ceil_div<long long>(n_items_cap > 0 ? n_items_cap : 1, groups_per_block);
  const long long cap = grid_cap();
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, (unsigned)K, (unsigned)col_chunks);
  gather_rows_kernel<VEC, LPR, NV, UNROLL><<<grid, kThreads, 0, st>>>(a);
  SG_LAUNCHED("gather_rows_kernel");
  if (a.hdr && n_long_cap > 0) {
    long long cb = ceil_div<long long>(n_long_cap, groups_per_block);
    if (cb > cap) cb = cap;
    dim3 cgrid((unsigned)cb, (unsigned)K, (unsigned)col_chunks);
    combine_partials_kernel<VEC, LPR, NV><<<cgrid, kThreads, 0, st>>>(a);
    SG_LAUNCHED("combine_partials_kernel");
  }
  return SG_OK;
}

template <int VEC>
static int dispatch_lpr(const GatherArgs &a, int K, int n_items_cap, int n_long_cap, cudaStream_t st) {
  const int vecs = ceil_div(a.F, VEC);
  if (vecs <= 1) return launch_gather<VEC, 1, 1>(a, K, n_items_cap, n_long_cap, st);
  if (vecs <= 2) return launch_gather<VEC, 2, 1>(a, K, n_items_cap, n_long_cap, st);
  if (vecs <= 4) return launch_gather<VEC, 4, 1>(a, K, n_items_cap, n_long_cap, st);
  if (vecs <= 8) return launch_gather<VEC, 8, 1>(a, K, n_items_cap, n_long_cap, st);
  if (vecs <= 16) return launch_gather<VEC, 16, 1>(a, K, n_items_cap, n_long_cap, st);
  if (vecs <= 32) return launch_gather<VEC, 32, 1>(a, K, n_items_cap, n_long_cap, st);
  if (vecs <= 64) return launch_gather<VEC, 32, 2>(a, K, n_items_cap, n_long_cap, st);
  return launch_gather<VEC, 32, 4>(a, K, n_items_cap, n_long_cap, st);  // wider rows loop over blockIdx.z
}

static bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// Threads per block of the fast gather launches.  Small blocks: (1) work items differ in length, so a block's slot
// is held until its slowest lane group is done — 64-thread blocks waste less of it (sum of the four ML-10M launches
// 0.790 ms at 256 threads, 0.757 at 128, 0.748 at 64; tools/sweep_gather.py); (2) a 64-thread block needs 3 - 4 K
// registers and fits next to a resident GEMM CTA (53.7 K of the SM's 64 K), so the other direction's gather makes
// progress while a transform runs instead of waiting for the SM to drain.
constexpr int kFastGatherThreads = 64;

template <int LPR, int NV, int UNROLL, int WMODE, bool WSUM, bool PLAIN = false, bool PEER = false>
static int launch_fast(const GatherArgs &a, int K, int n_items_cap, int n_long_cap, cudaStream_t st) {
  const int dv = dev_option(SG_DEV_GATHER_THREADS);
  const int kThreads = dv == 64 || dv == 128 || dv == 256 ? dv : kFastGatherThreads;
  const int groups_per_block = kThreads / LPR;
  long long blocks = ceil_div<long long>(n_items_cap > 0 ? n_items_cap : 1, groups_per_block);
  const long long cap = grid_cap() * (256 / kThreads);   // the same number of lane groups per SM whatever the block size
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, (unsigned)K, 1);
  gather_rows_fast_kernel<LPR, NV, UNROLL, WMODE, WSUM, PLAIN, PEER><<<grid, kThreads, 0, st>>>(a);
  SG_LAUNCHED("gather_rows_fast_kernel");
  if (a.hdr && n_long_cap > 0) {
    long long cb = ceil_div<long long>(n_long_cap, groups_per_block);
    if (cb > cap) cb = cap;
    dim3 cgrid((unsigned)cb, (unsigned)K, 1);
    combine_partials_kernel<4, LPR, NV, PEER><<<cgrid, kThreads, 0, st>>>(a);
    SG_LAUNCHED("combine_partials_kernel");
  }
  return SG_OK;
}

template <int LPR, int NV, int UNROLL>
static int dispatch_fast_mode(const GatherArgs &a, int K, int n_items_cap, int n_long_cap, cudaStream_t st) {
  if (a.peer_world > 0)   // rows scattered to their owners over NVLink (checked by run_gather: weights w[p], plain write)
    return launch_fast<LPR, NV, UNROLL, 1, false, true, true>(a, K, n_items_cap, n_long_cap, st);
  if (a.inv_len_indptr) return launch_fast<LPR, NV, UNROLL, 3, false>(a, K, n_items_cap, n_long_cap, st);
  if (!a.w) return launch_fast<LPR, NV, UNROLL, 0, false>(a, K, n_items_cap, n_long_cap, st);
  if (a.perm) return launch_fast<LPR, NV, UNROLL, 2, false>(a, K, n_items_cap, n_long_cap, st);
  const bool plain = !a.mean && a.req == SG_REQ_WRITE;
  if (a.wsum) return plain ? launch_fast<LPR, NV, UNROLL, 1, true, true>(a, K, n_items_cap, n_long_cap, st)
                           : launch_fast<LPR, NV, UNROLL, 1, true>(a, K, n_items_cap, n_long_cap, st);
  return plain ? launch_fast<LPR, NV, UNROLL, 1, false, true>(a, K, n_items_cap, n_long_cap, st)
               : launch_fast<LPR, NV, UNROLL, 1, false>(a, K, n_items_cap, n_long_cap, st);
}

template <bool WSUM>
static int launch_staged(const GatherArgs &a, int n_items_cap, int n_long_cap, cudaStream_t st) {
  long long blocks = ceil_div<long long>(n_items_cap > 0 ? n_items_cap : 1, 8);
  const long long cap = grid_cap();
  if (blocks > cap) blocks = cap;
  gather_rows_staged_kernel<WSUM><<<(unsigned)blocks, 256, 0, st>>>(a);
  SG_LAUNCHED("gather_rows_staged_kernel");
  if (n_long_cap > 0) {
    long long cb = ceil_div<long long>(n_long_cap, 16);
    if (cb > cap) cb = cap;
    combine_partials_kernel<4, 16, 1><<<dim3((unsigned)cb, 1, 1), 256, 0, st>>>(a);
    SG_LAUNCHED("combine_partials_kernel");
  }
  return SG_OK;
}

int run_gather(GatherArgs a, int K, int n_seg, int nnz, const void *plan, cudaStream_t st) {
  a.n_seg = n_seg;
  if (a.n_out_rows > 0) {  // multiply-high constants of seg / n_out_rows (seg_rel)
    const uint32_t d = (uint32_t)a.n_out_rows;
    int l = 0;
    while (l < 31 && (1u << l) < d) ++l;
    a.div_shift = l;
    a.div_magic = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  }
  int n_items_cap = n_seg, n_long_cap = 0;
  if (plan) {
    // header fields live on the device; size the grids from the host-side capacity bounds
    SG_REQUIRE(aligned(plan, 16), "plan buffer must be 16-byte aligned");
    SG_REQUIRE(a.plan_chunk > 0, "plan chunk must be given with a plan");
    const char *base = static_cast<const char *>(plan);
    a.hdr = reinterpret_cast<const PlanHeader *>(base);
    a.items = reinterpret_cast<const int4 *>(base + plan_off_items());
    a.longs = reinterpret_cast<const int4 *>(base + plan_off_longs(n_seg, nnz, a.plan_chunk));
    n_items_cap = (int)plan_cap_items(n_seg, nnz, a.plan_chunk);
    n_long_cap = (int)plan_cap_long(nnz, a.plan_chunk);
    if (n_long_cap > 0) SG_REQUIRE(a.partial, "a plan with split segments needs a partial-row scratch buffer");
  } else {
    a.hdr = nullptr; a.items = nullptr; a.longs = nullptr; a.partial = nullptr; a.partial_wsum = nullptr;
  }
  if (n_seg == 0 || a.F == 0) return SG_OK;
  if (a.peer_world > 0) {
    SG_REQUIRE(a.peer_world <= SG_MAX_PEERS + 1 && K == 1 && a.n_out_rows == n_seg && a.w && !a.perm && !a.inv_len_indptr &&
                   !a.wsum && !a.out_lo && !a.mean && a.req == SG_REQ_WRITE && a.ld_src == a.F && a.ld_out == a.F &&
                   (a.F == 16 || a.F == 32 || a.F == 64 || a.F == 128) && aligned(a.src, 16),
               "peer-scattered output needs the plain weighted fast path (F in 16/32/64/128, write, batch 1)");
    for (int q = 0; q < a.peer_world; ++q)
      SG_REQUIRE(a.peer_out[q] && aligned(a.peer_out[q], 16), "peer staging pointer %d null or unaligned", q);
    a.out = a.peer_out[0];   // only inspected for alignment below
  }
  const bool v4 = a.F % 4 == 0 && a.ld_src % 4 == 0 && a.ld_out % 4 == 0 && aligned(a.src, 16) && aligned(a.out, 16) &&
                  (!a.partial || aligned(a.partial, 16)) && a.src_batch_stride % 4 == 0 && a.out_batch_stride % 4 == 0 &&
                  a.partial_batch_stride % 4 == 0;
  const bool v2 = a.F % 2 == 0 && a.ld_src % 2 == 0 && a.ld_out % 2 == 0 && aligned(a.src, 8) && aligned(a.out, 8) &&
                  (!a.partial || aligned(a.partial, 8)) && a.src_batch_stride % 2 == 0 && a.out_batch_stride % 2 == 0 &&
                  a.partial_batch_stride % 2 == 0;
  if (v4 && a.ld_src == a.F && (!a.wsum || (a.w && !a.perm && !a.inv_len_indptr))) {  // exact-width fast path
    switch (a.F) {
      case 16: return dispatch_fast_mode<4, 1, 8>(a, K, n_items_cap, n_long_cap, st);
      case 32: return dispatch_fast_mode<8, 1, 8>(a, K, n_items_cap, n_long_cap, st);
      case 64:
        if (dev_option(SG_DEV_GATHER_VARIANT) == 1 && a.hdr && K == 1 && a.w && !a.perm && !a.inv_len_indptr && !a.mean &&
            a.req == SG_REQ_WRITE && nnz >= 16LL * n_items_cap && aligned(a.idx, 16) && aligned(a.w, 16))
          return a.wsum ? launch_staged<true>(a, n_items_cap, n_long_cap, st) : launch_staged<false>(a, n_items_cap, n_long_cap, st);
        // short segments (fewer than 16 edges per work item on average, the user side of a rating graph):
        // batches of 4 — as fast as batches of 8 there (0.263 ms both) with 46 instead of 64 registers
        if (nnz < 16LL * n_items_cap) return dispatch_fast_mode<16, 1, 4>(a, K, n_items_cap, n_long_cap, st);
        return dispatch_fast_mode<16, 1, 8>(a, K, n_items_cap, n_long_cap, st);
      case 128: return dispatch_fast_mode<32, 1, 8>(a, K, n_items_cap, n_long_cap, st);
      default: break;
    }
  }
  SG_REQUIRE(!a.out_lo, "pre-split (hi/lo) output needs 16-byte aligned rows of 16, 32, 64 or 128 floats");
  if (v4) return dispatch_lpr<4>(a, K, n_items_cap, n_long_cap, st);
  if (v2) return dispatch_lpr<2>(a, K, n_items_cap, n_long_cap, st);
  return dispatch_lpr<1>(a, K, n_items_cap, n_long_cap, st);
}

}  // namespace sg
