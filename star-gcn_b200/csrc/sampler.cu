// Device-side neighbour sampling, per-rating-level split and support.
//
// Replaces the host path of CSRMat.sample_neighbors (mxgraph/graph.py:677-748):
//   GraphSampler::random_sample_fix_neighbor   GraphSampler/graph_sampler.cpp:742-779
//   uniform_choice_range (partial Fisher-Yates) graph_sampler.cpp:698-732
//   multi_link_split_by_value                   graph_sampler.cpp:277-312 (serial = order-defining)
//   get_support                                 graph_sampler.cpp:393-420
//   np.take x 3 + per-level np.take x 3         graph.py:725-745
// followed in the reference by 4 H2D copies per level per call (layers.py:366-377).  Here the
// whole graph stays resident and the output IS the relation-major CSR the fused aggregation
// kernel consumes (segment id = r * n_sel + i).
//
// Integer outputs are bit-exact against the reference for the full neighbourhood (k < 0 or
// k >= degree): dst_ind_ptr, sampled positions, split indices and per-level ind_ptrs, and the
// support values (IEEE division and square root, as the host code computes them).  For k smaller
// than the degree the reference itself depends on OpenMP scheduling (128 mt19937 engines indexed
// by thread id, graph_sampler.cpp:765,776); here a counter-based generator keyed by
// (seed, global row, draw) makes the sample independent of launch shape and of how rows are
// partitioned over GPUs.
#include "common.cuh"

namespace sg {

constexpr int kMaxFanout = 256;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// uniform integer in [0, range) from (seed, row, draw)
__device__ __forceinline__ uint32_t draw_uniform(uint64_t seed, uint32_t row, uint32_t draw, uint32_t range) {
  const uint64_t h = mix64(seed ^ mix64(((uint64_t)row << 32) | draw));
  return (uint32_t)(((h >> 32) * (uint64_t)range) >> 32);
}

__global__ void __launch_bounds__(256) sample_count_kernel(int32_t *__restrict__ counts, const int32_t *__restrict__ src_indptr,
                                                           const int32_t *__restrict__ sel, int n_sel, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sel) return;
  const int row = sel ? __ldg(sel + i) : i;
  const int deg = __ldg(src_indptr + row + 1) - __ldg(src_indptr + row);
  counts[i] = k < 0 ? deg : min(k, deg);
}

// one warp per selected row.  Full neighbourhood: positions in order.  Otherwise lane 0 runs the
// partial Fisher-Yates of uniform_choice_range with the swapped entries kept in a small table
// (the reference keeps them in a hash map), k <= kMaxFanout.
__global__ void __launch_bounds__(256) sample_fill_kernel(int32_t *__restrict__ sampled, const int32_t *__restrict__ dst_indptr,
                                                          const int32_t *__restrict__ src_indptr,
                                                          const int32_t *__restrict__ sel, int n_sel, uint64_t seed) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n_sel; i += n_warps) {
    const int row = sel ? __ldg(sel + i) : i;
    const int p_begin = __ldg(src_indptr + row), p_end = __ldg(src_indptr + row + 1);
    const int shift = __ldg(dst_indptr + i);
    const int cnt = __ldg(dst_indptr + i + 1) - shift;
    const int deg = p_end - p_begin;
    if (cnt == deg) {
      for (int j = lane; j < deg; j += 32) sampled[shift + j] = p_begin + j;
    } else if (lane == 0) {
      int key[kMaxFanout], val[kMaxFanout];  // pool: virtual array entries that differ from identity
      int n_pool = 0;
      for (int lower = 0; lower < cnt; ++lower) {
        const int s = lower + (int)draw_uniform(seed, (uint32_t)row, (uint32_t)lower, (uint32_t)(deg - lower));
        int v_s = s, v_l = lower, at_s = -1;
        for (int t = 0; t < n_pool; ++t) {
          if (key[t] == s) { v_s = val[t]; at_s = t; }
          if (key[t] == lower) v_l = val[t];
        }
        sampled[shift + lower] = v_s + p_begin;
        if (at_s >= 0) val[at_s] = v_l;
        else { key[n_pool] = s; val[n_pool] = v_l; ++n_pool; }
      }
    }
  }
}

// support[j] = sqrt(1 / d_row / d_col)  |  1 / d_row     (graph_sampler.cpp:405-417, IEEE-exact)
__global__ void __launch_bounds__(256) support_kernel(float *__restrict__ support, const int32_t *__restrict__ row_deg,
                                                      const int32_t *__restrict__ col_deg,
                                                      const int32_t *__restrict__ indptr,
                                                      const int32_t *__restrict__ end_points, int n_rows, int symm) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n_rows; i += n_warps) {
    const int r_deg = __ldg(row_deg + i);
    for (int j = __ldg(indptr + i) + lane; j < __ldg(indptr + i + 1); j += 32) {
      float s = 0.f;
      if (symm) {
        const int c_deg = __ldg(col_deg + __ldg(end_points + j));
        if (r_deg != 0 && c_deg != 0) s = __fsqrt_rn(__fdiv_rn(__fdiv_rn(1.0f, (float)r_deg), (float)c_deg));
      } else if (r_deg != 0) {
        s = __fdiv_rn(1.0f, (float)r_deg);
      }
      support[j] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Static full-graph plans (stargcn_b200.static_step): instead of rebuilding both CSR directions without the batch
// edges every iteration (remove_edges, graph.py:952-974) the WHOLE training graph stays in one relation-major plan
// and only its edge weights change: a removed edge gets weight 0 (fma(0, x, acc) == acc: bit-neutral), a kept
// edge the normalisation 1/sqrt(d_row d_col) with the degrees of the graph WITHOUT the removed edges — evaluated
// exactly as get_support does (graph_sampler.cpp:393-420).
//   keep[p]          1 / 0 per position of the base CSR (sg_remove_edges_count's marking pass)
//   new_row_ptr      prefix sums of the kept edges per row (same call) -> degrees after removal
//   new_col_ptr      the same of the REVERSE direction's matrix (its rows are this matrix's columns)
//   split_index[q]   base position of plan position q,  plan_row[q] its row,  plan_col[q] its column
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) masked_support_kernel(float *__restrict__ support, const int32_t *__restrict__ keep,
                                                             const int32_t *__restrict__ new_row_ptr,
                                                             const int32_t *__restrict__ new_col_ptr,
                                                             const int32_t *__restrict__ split_index,
                                                             const int32_t *__restrict__ plan_row,
                                                             const int32_t *__restrict__ plan_col, int nnz, int symm) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += gridDim.x * blockDim.x) {
    float s = 0.f;
    if (__ldg(keep + __ldg(split_index + q)) != 0) {
      const int i = __ldg(plan_row + q);
      const int r_deg = __ldg(new_row_ptr + i + 1) - __ldg(new_row_ptr + i);
      if (symm) {
        const int j = __ldg(plan_col + q);
        const int c_deg = __ldg(new_col_ptr + j + 1) - __ldg(new_col_ptr + j);
        if (r_deg != 0 && c_deg != 0) s = __fsqrt_rn(__fdiv_rn(__fdiv_rn(1.0f, (float)r_deg), (float)c_deg));
      } else if (r_deg != 0) {
        s = __fdiv_rn(1.0f, (float)r_deg);
      }
    }
    support[q] = s;
  }
}

__device__ __forceinline__ int level_of(float v, const float *__restrict__ possible, int R) {
  int lvl = -1;
  for (int r = 0; r < R; ++r)
    if (__ldg(possible + r) == v) lvl = r;
  return lvl;
}

// counts[r * n_sel + i] = edges of sampled row i whose value is possible[r]; one warp per row,
// lane r keeps the running count of level r (R <= 32)
__global__ void __launch_bounds__(256) level_count_kernel(int32_t *__restrict__ counts, int32_t *__restrict__ bad,
                                                          const float *__restrict__ values,
                                                          const int32_t *__restrict__ sampled,
                                                          const int32_t *__restrict__ dst_indptr,
                                                          const float *__restrict__ possible, int R, int n_sel) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n_sel; i += n_warps) {
    const int lo = __ldg(dst_indptr + i), hi = __ldg(dst_indptr + i + 1);
    int mine = 0;
    for (int base = lo; base < hi; base += 32) {
      const int e = base + lane;
      int lvl = -2;
      if (e < hi) {
        lvl = level_of(__ldg(values + __ldg(sampled + e)), possible, R);
        if (lvl < 0) atomicOr(bad, 1);  // value outside the multi-link set (reference: ASSERT)
      }
      for (int r = 0; r < R; ++r) {
        const int c = __popc(__ballot_sync(0xffffffffu, lvl == r));
        if (lane == r) mine += c;
      }
    }
    if (lane < R) counts[(long long)lane * n_sel + i] = mine;
  }
}

// stable per-level scatter: position inside (level, row) follows the order of the sampled row
__global__ void __launch_bounds__(256) level_scatter_kernel(int32_t *__restrict__ split_index, int32_t *__restrict__ ep_cat,
                                                            float *__restrict__ sup_cat, float *__restrict__ val_cat,
                                                            const int32_t *__restrict__ cat_indptr,
                                                            const float *__restrict__ values,
                                                            const int32_t *__restrict__ end_points,
                                                            const float *__restrict__ support,
                                                            const int32_t *__restrict__ sampled,
                                                            const int32_t *__restrict__ dst_indptr,
                                                            const float *__restrict__ possible, int R, int n_sel) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (int i = warp; i < n_sel; i += n_warps) {
    const int lo = __ldg(dst_indptr + i), hi = __ldg(dst_indptr + i + 1);
    // lane r: next free slot of (level r, row i)
    int next = lane < R ? __ldg(cat_indptr + (long long)lane * n_sel + i) : 0;
    for (int base = lo; base < hi; base += 32) {
      const int e = base + lane;
      int lvl = -2, src = 0;
      if (e < hi) {
        src = __ldg(sampled + e);
        lvl = level_of(__ldg(values + src), possible, R);
      }
      for (int r = 0; r < R; ++r) {
        const unsigned m = __ballot_sync(0xffffffffu, lvl == r);
        const int first = __shfl_sync(0xffffffffu, next, r);
        if (lvl == r) {
          const int pos = first + __popc(m & lt);
          if (split_index) split_index[pos] = e;
          ep_cat[pos] = __ldg(end_points + src);
          if (sup_cat) sup_cat[pos] = __ldg(support + src);
          if (val_cat) val_cat[pos] = __ldg(values + src);
        }
        if (lane == r) next += __popc(m);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// batch-edge removal: remove_edges, GraphSampler/graph_sampler.cpp:154-201 — the
// reference rebuilds both CSR directions on ONE host thread every training iteration
// (experiments/STAR-GCN.py:595-600).  Here: mark (one warp per removal pair scans its row),
// count kept edges per row, scan, stable compaction (one warp per row, ballot).  Order inside
// rows is preserved, every copy of a listed (row, col) pair is dropped: bit-exact outputs.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mark_removed_kernel(int32_t *__restrict__ keep, const int32_t *__restrict__ indptr,
                                                           const int32_t *__restrict__ end_points,
                                                           const int32_t *__restrict__ rm_rows,
                                                           const int32_t *__restrict__ rm_cols, int n_rows, int n_rm) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int e = warp; e < n_rm; e += n_warps) {
    const int r = __ldg(rm_rows + e), c = __ldg(rm_cols + e);
    if (r < 0 || r >= n_rows) continue;
    for (int p = __ldg(indptr + r) + lane; p < __ldg(indptr + r + 1); p += 32)
      if (__ldg(end_points + p) == c) keep[p] = 0;
  }
}

__global__ void __launch_bounds__(256) count_kept_kernel(int32_t *__restrict__ counts, const int32_t *__restrict__ keep,
                                                         const int32_t *__restrict__ indptr, int n_rows) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n_rows; i += n_warps) {
    int c = 0;
    for (int p = __ldg(indptr + i) + lane; p < __ldg(indptr + i + 1); p += 32) c += __ldg(keep + p) != 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0) counts[i] = c;
  }
}

__global__ void __launch_bounds__(256) compact_rows_kernel(int32_t *__restrict__ dst_ep, float *__restrict__ dst_val,
                                                           const int32_t *__restrict__ dst_indptr,
                                                           const int32_t *__restrict__ keep,
                                                           const int32_t *__restrict__ end_points,
                                                           const float *__restrict__ values,
                                                           const int32_t *__restrict__ indptr, int n_rows) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (int i = warp; i < n_rows; i += n_warps) {
    const int lo = __ldg(indptr + i), hi = __ldg(indptr + i + 1);
    int out = __ldg(dst_indptr + i);
    for (int base = lo; base < hi; base += 32) {
      const int p = base + lane;
      const bool k = p < hi && __ldg(keep + p) != 0;
      const unsigned m = __ballot_sync(0xffffffffu, k);
      if (k) {
        const int o = out + __popc(m & lt);
        dst_ep[o] = __ldg(end_points + p);
        if (dst_val) dst_val[o] = __ldg(values + p);
      }
      out += __popc(m);
    }
  }
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t *__restrict__ out, int v, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = v;
}

// integer histogram (degrees of the column side); integer atomics: the result is order-independent
__global__ void __launch_bounds__(256) bincount_kernel(int32_t *__restrict__ counts, const int32_t *__restrict__ idx, int n,
                                                       int n_bins) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int b = __ldg(idx + i);
    if (b >= 0 && b < n_bins) atomicAdd(counts + b, 1);
  }
}

static inline int grid_w(long long n_warps_needed) {
  long long g = ceil_div<long long>(n_warps_needed > 0 ? n_warps_needed : 1, 8);
  long long cap = (long long)num_sms() * 16;
  return (int)(g < cap ? g : cap);
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_csr_support(float *support, const int32_t *row_degrees, const int32_t *col_degrees, const int32_t *indptr,
                   const int32_t *end_points, int n_rows, int nnz, int symm, sg_stream_t stream) {
  SG_REQUIRE(n_rows >= 0 && nnz >= 0, "sg_csr_support: negative size");
  if (n_rows == 0 || nnz == 0) return SG_OK;
  SG_REQUIRE(support && row_degrees && indptr && (!symm || (col_degrees && end_points)), "sg_csr_support: null pointer");
  support_kernel<<<grid_w(n_rows), 256, 0, (cudaStream_t)stream>>>(support, row_degrees, col_degrees, indptr, end_points, n_rows, symm);
  SG_LAUNCHED("support_kernel");
  return SG_OK;
}

size_t sg_sampler_ws_bytes(int n_sel, int R) {
  const int n = (n_sel > 0 ? n_sel : 1) * (R > 0 ? R : 1) + 1;
  return align_up((size_t)n * sizeof(int32_t), 64) + scan_ws_bytes(n) + 128;
}

int sg_sample_neighbors_count(int32_t *dst_indptr, const int32_t *src_indptr, const int32_t *sel, int n_sel,
                              int neighbor_num, void *ws, sg_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_REQUIRE(n_sel >= 0, "sg_sample_neighbors_count: negative size");
  SG_REQUIRE(dst_indptr && ws && (n_sel == 0 || src_indptr), "sg_sample_neighbors_count: null pointer");
  SG_REQUIRE(neighbor_num < 0 || neighbor_num <= kMaxFanout, "sg_sample_neighbors_count: fan-out %d exceeds the supported maximum %d", neighbor_num, kMaxFanout);
  if (n_sel == 0) { SG_CUDA(cudaMemsetAsync(dst_indptr, 0, sizeof(int32_t), st)); return SG_OK; }
  sample_count_kernel<<<ceil_div(n_sel, 256), 256, 0, st>>>(dst_indptr, src_indptr, sel, n_sel, neighbor_num);
  SG_LAUNCHED("sample_count_kernel");
  // exclusive scan in place; the total lands in dst_indptr[n_sel]
  return exclusive_scan_i32(dst_indptr, dst_indptr, n_sel, dst_indptr + n_sel, ws, st);
}

int sg_sample_neighbors_fill(int32_t *sampled, const int32_t *dst_indptr, const int32_t *src_indptr, const int32_t *sel,
                             int n_sel, unsigned long long seed, sg_stream_t stream) {
  SG_REQUIRE(n_sel >= 0, "sg_sample_neighbors_fill: negative size");
  if (n_sel == 0) return SG_OK;
  SG_REQUIRE(sampled && dst_indptr && src_indptr, "sg_sample_neighbors_fill: null pointer");
  sample_fill_kernel<<<grid_w(n_sel), 256, 0, (cudaStream_t)stream>>>(sampled, dst_indptr, src_indptr, sel, n_sel, (uint64_t)seed);
  SG_LAUNCHED("sample_fill_kernel");
  return SG_OK;
}

int sg_multilink_split(int32_t *cat_indptr, int32_t *split_index, int32_t *ep_cat, float *sup_cat, float *val_cat,
                       int32_t *bad_flag, const float *values, const int32_t *end_points, const float *support,
                       const int32_t *sampled, const int32_t *dst_indptr, const float *possible_values, int R, int n_sel,
                       void *ws, sg_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_REQUIRE(R > 0 && R <= 32 && n_sel >= 0, "sg_multilink_split: bad sizes (1 <= R <= 32)");
  SG_REQUIRE((long long)R * n_sel < (1LL << 31) - 1, "sg_multilink_split: R * n_sel overflows int32");
  SG_REQUIRE(cat_indptr && ws && bad_flag, "sg_multilink_split: null pointer");
  SG_CUDA(cudaMemsetAsync(bad_flag, 0, sizeof(int32_t), st));
  if (n_sel == 0) { SG_CUDA(cudaMemsetAsync(cat_indptr, 0, sizeof(int32_t), st)); return SG_OK; }
  SG_REQUIRE(ep_cat && values && end_points && sampled && dst_indptr && possible_values && (!sup_cat || support),
             "sg_multilink_split: null pointer");
  const int n = R * n_sel;
  level_count_kernel<<<grid_w(n_sel), 256, 0, st>>>(cat_indptr, bad_flag, values, sampled, dst_indptr, possible_values, R, n_sel);
  SG_LAUNCHED("level_count_kernel");
  int rc = exclusive_scan_i32(cat_indptr, cat_indptr, n, cat_indptr + n, ws, st);
  if (rc != SG_OK) return rc;
  level_scatter_kernel<<<grid_w(n_sel), 256, 0, st>>>(split_index, ep_cat, sup_cat, val_cat, cat_indptr, values, end_points, support,
                                                      sampled, dst_indptr, possible_values, R, n_sel);
  SG_LAUNCHED("level_scatter_kernel");
  return SG_OK;
}

size_t sg_remove_edges_ws_bytes(int n_rows, int nnz) {
  return align_up((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t), 64) + scan_ws_bytes(n_rows > 0 ? n_rows : 1) + 128;
}

int sg_remove_edges_count(int32_t *dst_indptr, const int32_t *indptr, const int32_t *end_points, const int32_t *rm_rows,
                          const int32_t *rm_cols, int n_rows, int nnz, int n_rm, void *ws, sg_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_REQUIRE(n_rows >= 0 && nnz >= 0 && n_rm >= 0, "sg_remove_edges_count: negative size");
  SG_REQUIRE(dst_indptr && ws && (n_rows == 0 || indptr), "sg_remove_edges_count: null pointer");
  if (n_rows == 0) { SG_CUDA(cudaMemsetAsync(dst_indptr, 0, sizeof(int32_t), st)); return SG_OK; }
  int32_t *keep = static_cast<int32_t *>(ws);
  void *scan_ws = static_cast<char *>(ws) + align_up((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t), 64);
  if (nnz > 0) {
    SG_REQUIRE(end_points && (n_rm == 0 || (rm_rows && rm_cols)), "sg_remove_edges_count: null pointer");
    fill_i32_kernel<<<grid_w(ceil_div(nnz, 32)), 256, 0, st>>>(keep, 1, nnz);
    SG_LAUNCHED("fill_i32_kernel");
    if (n_rm > 0) {
      mark_removed_kernel<<<grid_w(n_rm), 256, 0, st>>>(keep, indptr, end_points, rm_rows, rm_cols, n_rows, n_rm);
      SG_LAUNCHED("mark_removed_kernel");
    }
  }
  count_kept_kernel<<<grid_w(n_rows), 256, 0, st>>>(dst_indptr, keep, indptr, n_rows);
  SG_LAUNCHED("count_kept_kernel");
  return exclusive_scan_i32(dst_indptr, dst_indptr, n_rows, dst_indptr + n_rows, scan_ws, st);
}

int sg_remove_edges_fill(int32_t *dst_end_points, float *dst_values, const int32_t *dst_indptr, const int32_t *indptr,
                         const int32_t *end_points, const float *values, int n_rows, int nnz, const void *ws,
                         sg_stream_t stream) {
  SG_REQUIRE(n_rows >= 0 && nnz >= 0, "sg_remove_edges_fill: negative size");
  if (n_rows == 0 || nnz == 0) return SG_OK;
  SG_REQUIRE(dst_end_points && dst_indptr && indptr && end_points && ws && (!dst_values || values), "sg_remove_edges_fill: null pointer");
  compact_rows_kernel<<<grid_w(n_rows), 256, 0, (cudaStream_t)stream>>>(dst_end_points, dst_values, dst_indptr,
                                                                        static_cast<const int32_t *>(ws), end_points, values, indptr, n_rows);
  SG_LAUNCHED("compact_rows_kernel");
  return SG_OK;
}

int sg_masked_support(float *support, const int32_t *keep, const int32_t *new_row_ptr, const int32_t *new_col_ptr,
                      const int32_t *split_index, const int32_t *plan_row, const int32_t *plan_col, int nnz, int symm,
                      sg_stream_t stream) {
  SG_REQUIRE(nnz >= 0, "sg_masked_support: negative size");
  if (nnz == 0) return SG_OK;
  SG_REQUIRE(support && keep && new_row_ptr && split_index && plan_row && (!symm || (new_col_ptr && plan_col)),
             "sg_masked_support: null pointer");
  masked_support_kernel<<<grid_w(ceil_div(nnz, 32)), 256, 0, (cudaStream_t)stream>>>(support, keep, new_row_ptr, new_col_ptr,
                                                                                    split_index, plan_row, plan_col, nnz, symm);
  SG_LAUNCHED("masked_support_kernel");
  return SG_OK;
}

int sg_bincount(int32_t *counts, const int32_t *idx, int n, int n_bins, sg_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_REQUIRE(n >= 0 && n_bins >= 0, "sg_bincount: negative size");
  if (n_bins == 0) return SG_OK;
  SG_REQUIRE(counts && (n == 0 || idx), "sg_bincount: null pointer");
  SG_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n_bins, st));
  if (n == 0) return SG_OK;
  bincount_kernel<<<grid_w(ceil_div(n, 32)), 256, 0, st>>>(counts, idx, n, n_bins);
  SG_LAUNCHED("bincount_kernel");
  return SG_OK;
}

}  // extern "C"
