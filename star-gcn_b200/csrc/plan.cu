// Segment bookkeeping: error state, int32 scan, segment ids, stable CSR transpose and the
// work-item schedule ("plan") the gather kernels walk.  All integer work — bit-exact against
// oracle/seg_ops_oracle.c (orc_seg_ids, orc_csr_transpose) and the reference's
// gen_row_indices_by_indptr (GraphSampler/graph_sampler.cpp:378-391).
#include <atomic>
#include <cstdarg>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace sg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};  // process-wide: autograd runs backward on its own thread

static std::atomic<int> g_dev_options[SG_DEV_COUNT];

char *err_buf() { return g_err; }
int dev_option(int which) { return which >= 0 && which < SG_DEV_COUNT ? g_dev_options[which].load(std::memory_order_relaxed) : 0; }
void count_launch(int n) { g_launches += n; }

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ------------------------------------------------------------------------------------------
// exclusive scan (3 launches; n <= 2^31).  2048 elements per 256-thread block.
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 8;
constexpr int kScanTile = kScanThreads * kScanPerThread;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// inclusive scan of one value per thread across a 256-thread block; returns inclusive value,
// *block_total gets the sum.
__device__ __forceinline__ int block_incl_scan(int v, int *block_total) {
  __shared__ int warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int s = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    s = warp_incl_scan(s, lane);
    if (lane < kScanThreads / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  if (wid > 0) inc += warp_sums[wid - 1];
  *block_total = warp_sums[kScanThreads / 32 - 1];
  __syncthreads();
  return inc;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int32_t *__restrict__ in, int n,
                                                               int32_t *__restrict__ tile_sums) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanPerThread;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i)
    if (base + i < n) s += in[base + i];
  int total;
  block_incl_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_of_tile_sums(int32_t *tile_sums, int n_tiles,
                                                                  int32_t *total_out) {
  int carry = 0;
  for (int base = 0; base < n_tiles; base += kScanThreads) {
    int i = base + threadIdx.x;
    int v = i < n_tiles ? tile_sums[i] : 0;
    int total;
    int inc = block_incl_scan(v, &total);
    if (i < n_tiles) tile_sums[i] = carry + inc - v;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply(int32_t *out, const int32_t *in, int n,
                                                           const int32_t *__restrict__ tile_offs) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanPerThread;
  int v[kScanPerThread];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i) {
    v[i] = base + i < n ? in[base + i] : 0;
    s += v[i];
  }
  int total;
  int run = block_incl_scan(s, &total) - s + tile_offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

size_t scan_ws_bytes(int n) { return align_up((size_t)(ceil_div(n > 0 ? n : 1, kScanTile)) * sizeof(int32_t), 64); }

int exclusive_scan_i32(int32_t *out, const int32_t *in, int n, int32_t *total, void *ws, cudaStream_t st) {
  if (n <= 0) {
    if (total) SG_CUDA(cudaMemsetAsync(total, 0, sizeof(int32_t), st));
    return SG_OK;
  }
  int32_t *tile_sums = static_cast<int32_t *>(ws);
  const int tiles = ceil_div(n, kScanTile);
  scan_tile_sums<<<tiles, kScanThreads, 0, st>>>(in, n, tile_sums);
  SG_LAUNCHED("scan_tile_sums");
  scan_of_tile_sums<<<1, kScanThreads, 0, st>>>(tile_sums, tiles, total);
  SG_LAUNCHED("scan_of_tile_sums");
  scan_apply<<<tiles, kScanThreads, 0, st>>>(out, in, n, tile_sums);
  SG_LAUNCHED("scan_apply");
  return SG_OK;
}

// ------------------------------------------------------------------------------------------
// segment ids: position p -> the segment s with indptr[s] <= p < indptr[s+1]
// (upper_bound over indptr, so runs of empty segments are skipped correctly)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int seg_of_pos(const int32_t *__restrict__ indptr, int n_seg, int p) {
  int lo = 0, hi = n_seg;  // find first s with indptr[s+1] > p
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(indptr + mid + 1) > p) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// A block owns a tile of kSegTile consecutive positions: two global searches find the first and last
// segment that overlap it, their indptr entries are staged in shared memory, and every position then
// searches that window (<= 12 steps in shared memory instead of ~20 dependent global loads: 92 -> ~20 us for
// 8 M positions over 700 k segments).  A window wider than the staging buffer (long runs of empty segments)
// falls back to the global search.  Same result as one seg_of_pos per position.
constexpr int kSegTile = 2048;
constexpr int kSegWin = 2560;  // staged indptr entries

__global__ void __launch_bounds__(256) seg_ids_kernel(int32_t *__restrict__ seg_ids,
                                                      const int32_t *__restrict__ indptr, int n_seg, int nnz) {
  __shared__ int32_t s_ptr[kSegWin];
  __shared__ int s_lohi[2];
  const int n_tiles = ceil_div(nnz, kSegTile);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int p0 = tile * kSegTile, p1 = min(p0 + kSegTile, nnz);
    if (threadIdx.x == 0) s_lohi[0] = seg_of_pos(indptr, n_seg, p0);
    if (threadIdx.x == 32) s_lohi[1] = seg_of_pos(indptr, n_seg, p1 - 1);
    __syncthreads();
    const int lo = s_lohi[0], cnt = s_lohi[1] - lo + 1;  // segments lo .. lo + cnt - 1 overlap the tile
    const bool staged = cnt + 1 <= kSegWin;
    if (staged)
      for (int i = threadIdx.x; i <= cnt; i += blockDim.x) s_ptr[i] = __ldg(indptr + lo + i);
    __syncthreads();
    for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
      int seg;
      if (staged) {
        int l = 0, h = cnt - 1;  // first j with s_ptr[j + 1] > p; j = cnt - 1 always qualifies
        while (l < h) {
          const int m = (l + h) >> 1;
          if (s_ptr[m + 1] > p) h = m; else l = m + 1;
        }
        seg = lo + l;
      } else {
        seg = seg_of_pos(indptr, n_seg, p);
      }
      seg_ids[p] = seg;
    }
    __syncthreads();  // the next tile overwrites the window
  }
}

// ------------------------------------------------------------------------------------------
// plan build in three launches: per-tile sums of (items, long segments, partial slots) straight from
// indptr, one single-block scan of the tile sums (three channels), and a fill pass that redoes the block
// scan and writes the work items.  (The first version materialised three count arrays and ran a
// 3-launch scan on each: 11 launches per schedule, two schedules per plan, rebuilt with every plan.)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void plan_counts(int len, int chunk, int &n_item, int &n_long, int &n_part) {
  const bool is_long = len > chunk;
  n_item = is_long ? ceil_div(len, chunk) : 1;
  n_long = is_long ? 1 : 0;
  n_part = is_long ? n_item : 0;
}

__global__ void __launch_bounds__(kScanThreads) plan_tile_sums(const int32_t *__restrict__ indptr, int n_seg, int chunk,
                                                               int32_t *__restrict__ tile_sums, int n_tiles) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanPerThread;
  int s_item = 0, s_long = 0, s_part = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i) {
    const int s = base + i;
    if (s < n_seg) {
      int a, b, c;
      plan_counts(__ldg(indptr + s + 1) - __ldg(indptr + s), chunk, a, b, c);
      s_item += a; s_long += b; s_part += c;
    }
  }
  int t_item, t_long, t_part;
  block_incl_scan(s_item, &t_item);
  block_incl_scan(s_long, &t_long);
  block_incl_scan(s_part, &t_part);
  if (threadIdx.x == 0) {
    tile_sums[blockIdx.x] = t_item;
    tile_sums[n_tiles + blockIdx.x] = t_long;
    tile_sums[2 * n_tiles + blockIdx.x] = t_part;
  }
}

// exclusive scan of the three channels of tile sums in place; totals -> hdr->n_items / n_long / n_partials
__global__ void __launch_bounds__(kScanThreads) plan_scan_tiles(int32_t *tile_sums, int n_tiles, PlanHeader *hdr) {
  for (int ch = 0; ch < 3; ++ch) {
    int32_t *t = tile_sums + ch * n_tiles;
    int carry = 0;
    for (int base = 0; base < n_tiles; base += kScanThreads) {
      const int i = base + threadIdx.x;
      const int v = i < n_tiles ? t[i] : 0;
      int total;
      const int inc = block_incl_scan(v, &total);
      if (i < n_tiles) t[i] = carry + inc - v;
      carry += total;
    }
    if (threadIdx.x == 0) {
      if (ch == 0) hdr->n_items = carry;
      else if (ch == 1) hdr->n_long = carry;
      else hdr->n_partials = carry;
    }
  }
}

__global__ void __launch_bounds__(kScanThreads) plan_fill_fused(PlanHeader *hdr, int4 *__restrict__ items, int4 *__restrict__ longs,
                                                                const int32_t *__restrict__ indptr, int n_seg, int nnz, int chunk,
                                                                const int32_t *__restrict__ tile_offs, int n_tiles,
                                                                int cap_items, int cap_long) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hdr->chunk = chunk; hdr->n_seg = n_seg; hdr->nnz = nnz; hdr->cap_items = cap_items; hdr->cap_long = cap_long;
  }
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanPerThread;
  int beg[kScanPerThread + 1];
#pragma unroll
  for (int i = 0; i <= kScanPerThread; ++i) beg[i] = base + i <= n_seg ? __ldg(indptr + base + i) : 0;
  int s_item = 0, s_long = 0, s_part = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i) {
    if (base + i < n_seg) {
      int a, b, c;
      plan_counts(beg[i + 1] - beg[i], chunk, a, b, c);
      s_item += a; s_long += b; s_part += c;
    }
  }
  int total;
  int o_item = block_incl_scan(s_item, &total) - s_item + tile_offs[blockIdx.x];
  int o_long = block_incl_scan(s_long, &total) - s_long + tile_offs[n_tiles + blockIdx.x];
  int o_part = block_incl_scan(s_part, &total) - s_part + tile_offs[2 * n_tiles + blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i) {
    const int s = base + i;
    if (s >= n_seg) break;
    const int b0 = beg[i], e0 = beg[i + 1], len = e0 - b0;
    if (len <= chunk) {
      items[o_item] = make_int4(b0, e0, s, -1);
      o_item += 1;
    } else {
      const int nch = ceil_div(len, chunk);
      for (int c = 0; c < nch; ++c) {
        const int b = b0 + c * chunk;
        items[o_item + c] = make_int4(b, min(b + chunk, e0), s, o_part + c);
      }
      longs[o_long] = make_int4(s, o_part, nch, 0);
      o_item += nch; o_long += 1; o_part += nch;
    }
  }
}

// ------------------------------------------------------------------------------------------
// transpose helpers
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) iota_kernel(int32_t *__restrict__ out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i;
}

// t_indptr[n] = number of sorted keys < n  (lower_bound), n in [0, n_nb]
__global__ void __launch_bounds__(256) indptr_from_sorted(int32_t *__restrict__ t_indptr,
                                                          const int32_t *__restrict__ keys, int nnz, int n_nb) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n > n_nb) return;
  int lo = 0, hi = nnz;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < n) lo = mid + 1; else hi = mid;
  }
  t_indptr[n] = lo;
}

__global__ void __launch_bounds__(256) gather_i32(int32_t *__restrict__ out, const int32_t *__restrict__ src,
                                                  const int32_t *__restrict__ idx, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = src[idx[i]];
}

__global__ void __launch_bounds__(256) multilink_finish(int32_t *__restrict__ t_src, float *__restrict__ t_w,
                                                        const int32_t *__restrict__ t_perm,
                                                        const int32_t *__restrict__ t_seg,
                                                        const float *__restrict__ support, int R, int n_dst, int nnz) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += gridDim.x * blockDim.x) {
    int s = t_seg[q];
    int r = s / n_dst, i = s - r * n_dst;
    t_src[q] = i * R + r;
    t_w[q] = support[t_perm[q]];
  }
}

// ------------------------------------------------------------------------------------------
// unique + inverse in FIRST-OCCURRENCE order: the serial, order-defining
// unique_inverse of GraphSampler/graph_sampler.h:510-534 that merge_nodes / gen_plan use to turn
// node ids into local row indices (mxgraph/graph.py:142-163, layers.py:308-334).
//   stable sort (value, position) -> run heads -> rank the runs by the position of their first
//   element -> unique[rank] = value, inverse[position] = rank.  Integer work, bit-exact.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) run_heads_kernel(int32_t *__restrict__ head, const int32_t *__restrict__ ks, int n) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    head[j] = (j == 0 || ks[j] != ks[j - 1]) ? 1 : 0;
}

// run_of[j] = index of the run sorted element j belongs to; first_pos[run] = original position of its head
__global__ void __launch_bounds__(256) run_first_pos_kernel(int32_t *__restrict__ first_pos, int32_t *__restrict__ run_iota,
                                                            const int32_t *__restrict__ head_excl,
                                                            const int32_t *__restrict__ ks, const int32_t *__restrict__ pos,
                                                            int n) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    if (j == 0 || ks[j] != ks[j - 1]) {
      const int u = head_excl[j];
      first_pos[u] = pos[j];
      run_iota[u] = u;
    }
  }
}

__global__ void __launch_bounds__(256) rank_scatter_kernel(int32_t *__restrict__ rank, const int32_t *__restrict__ order,
                                                           const int32_t *__restrict__ n_unique) {
  const int n = *n_unique;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) rank[order[k]] = k;
}

__global__ void __launch_bounds__(256) unique_write_kernel(int32_t *__restrict__ uniq, int32_t *__restrict__ inverse,
                                                           const int32_t *__restrict__ rank,
                                                           const int32_t *__restrict__ head_excl,
                                                           const int32_t *__restrict__ ks, const int32_t *__restrict__ pos,
                                                           int n) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const bool is_head = j == 0 || ks[j] != ks[j - 1];
    const int u = head_excl[j] + (is_head ? 0 : -1);  // exclusive count of heads before j; non-heads belong to the previous head
    const int r = rank[u];
    if (is_head) uniq[r] = ks[j];
    inverse[pos[j]] = r;
  }
}

static inline int grid_for(long long n, int threads = 256) {
  long long g = ceil_div<long long>(n > 0 ? n : 1, threads);
  long long cap = (long long)num_sms() * 32;
  return (int)(g < cap ? g : cap);
}

static inline int seg_ids_grid(int nnz) {  // one block per tile of positions, grid-stride beyond 32 blocks per SM
  const long long tiles = ceil_div<long long>(nnz > 0 ? nnz : 1, kSegTile);
  const long long cap = (long long)num_sms() * 32;
  return (int)(tiles < cap ? tiles : cap);
}

static int key_bits(int n_nb) {
  int bits = 1;
  while (bits < 31 && (1LL << bits) < (long long)n_nb) ++bits;
  return bits;
}

// ------------------------------------------------------------------------------------------
// Upload of many small host arrays in ONE launch: the per-level lists of a plan (3 * R arrays per direction,
// mxgraph/layers/layers.py:366-377 uploads each with its own nd.array call) sit in PINNED host memory, which a
// kernel can read directly (unified addressing): every thread copies one 16-byte unit of the flattened segment
// list, a warp request is 512 contiguous bytes of one host array.  For the small plans of ML-100k / Douban-sized
// graphs this replaces ~60 DMA set-ups (~10 us each) by one ~20 us kernel; large plans keep the copy engines.
// ------------------------------------------------------------------------------------------
constexpr int kUploadMaxSeg = 64;
struct UploadArgs {
  const uint32_t *src[kUploadMaxSeg];
  uint32_t *dst[kUploadMaxSeg];
  long long end[kUploadMaxSeg];   // cumulative counts of 16-byte units (4 words; the last unit of a segment may be partial)
  long long words[kUploadMaxSeg]; // 4-byte words of the segment
  int n;
};

// `end` counts UNITS of four 4-byte words per segment (the last unit of a segment may be partial): a thread reads one
// unit with ONE 16-byte load when the host array allows it (16-byte aligned base — pinned allocations are) — a warp
// request is then 512 contiguous bytes of host memory instead of 128, which is what the PCIe read path wants — and
// stores it with one 16-byte store when the device slice is aligned too, else word by word.
__global__ void __launch_bounds__(256) upload_segments_kernel(const __grid_constant__ UploadArgs a) {
  const long long total = a.end[a.n - 1];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int k = 0;
    while (t >= a.end[k]) ++k;            // n <= 64 segments: a short scan, uniform inside a warp except at seams
    const long long u = t - (k ? a.end[k - 1] : 0);
    const long long w0 = u * 4, nw = a.words[k];
    const uint32_t *src = a.src[k] + w0;
    uint32_t *dst = a.dst[k] + w0;
    if (w0 + 4 <= nw && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const uint4 v = *reinterpret_cast<const uint4 *>(src);
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        *reinterpret_cast<uint4 *>(dst) = v;
      } else {
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      }
    } else {
      for (int j = 0; j < 4 && w0 + j < nw; ++j) dst[j] = src[j];
    }
  }
}

}  // namespace sg

using namespace sg;

extern "C" {

const char *sg_last_error(void) { return sg::err_buf(); }
int sg_abi_version(void) { return 1; }
long long sg_launch_count(void) { return sg::g_launches.load(); }
void sg_launch_count_reset(void) { sg::g_launches.store(0); }
int sg_dev_option(int which, int value) {
  SG_REQUIRE(which >= 0 && which < sg::SG_DEV_COUNT, "sg_dev_option: unknown option %d", which);
  sg::g_dev_options[which].store(value);
  return SG_OK;
}

int sg_seg_ids(int32_t *seg_ids, const int32_t *indptr, int n_seg, int nnz, sg_stream_t stream) {
  SG_REQUIRE(n_seg >= 0 && nnz >= 0, "sg_seg_ids: negative size (n_seg=%d nnz=%d)", n_seg, nnz);
  if (nnz == 0) return SG_OK;
  SG_REQUIRE(seg_ids && indptr, "sg_seg_ids: null pointer");
  seg_ids_kernel<<<seg_ids_grid(nnz), 256, 0, (cudaStream_t)stream>>>(seg_ids, indptr, n_seg, nnz);
  SG_LAUNCHED("seg_ids_kernel");
  return SG_OK;
}

size_t sg_csr_transpose_ws_bytes(int n_seg, int n_nb, int nnz) {
  (void)n_seg;
  if (nnz <= 0) return 64;
  size_t cub_bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                                  (const int32_t *)nullptr, (int32_t *)nullptr, nnz, 0,
                                                  key_bits(n_nb));
  if (e != cudaSuccess) {
    sg::fail(SG_ERR_CUDA, "sg_csr_transpose_ws_bytes: CUB size query failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return 0;
  }
  // sorted keys + iota + seg ids + CUB scratch
  return 3 * align_up((size_t)nnz * sizeof(int32_t), 256) + align_up(cub_bytes, 256) + 256;
}

int sg_csr_transpose(int32_t *t_indptr, int32_t *t_perm, int32_t *t_seg, const int32_t *indices,
                     const int32_t *indptr, int n_seg, int n_nb, int nnz, void *ws, size_t ws_bytes,
                     sg_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_REQUIRE(n_seg >= 0 && n_nb >= 0 && nnz >= 0, "sg_csr_transpose: negative size");
  SG_REQUIRE(t_indptr, "sg_csr_transpose: null t_indptr");
  if (nnz == 0) {
    SG_CUDA(cudaMemsetAsync(t_indptr, 0, sizeof(int32_t) * (size_t)(n_nb + 1), st));
    return SG_OK;
  }
  SG_REQUIRE(t_perm && t_seg && indices && indptr && ws, "sg_csr_transpose: null pointer");
  const size_t need = sg_csr_transpose_ws_bytes(n_seg, n_nb, nnz);
  if (need == 0) return SG_ERR_CUDA;
  if (ws_bytes < need) return sg::fail(SG_ERR_WORKSPACE, "sg_csr_transpose: workspace %zu < %zu bytes", ws_bytes, need);
  const size_t stride = align_up((size_t)nnz * sizeof(int32_t), 256);
  char *base = static_cast<char *>(ws);
  int32_t *keys_sorted = reinterpret_cast<int32_t *>(base);
  int32_t *iota = reinterpret_cast<int32_t *>(base + stride);
  int32_t *seg_ids = reinterpret_cast<int32_t *>(base + 2 * stride);
  void *cub_ws = base + 3 * stride;
  size_t cub_bytes = ws_bytes - 3 * stride;

  iota_kernel<<<grid_for(nnz), 256, 0, st>>>(iota, nnz);
  SG_LAUNCHED("iota_kernel");
  seg_ids_kernel<<<seg_ids_grid(nnz), 256, 0, st>>>(seg_ids, indptr, n_seg, nnz);
  SG_LAUNCHED("seg_ids_kernel");
  // LSD radix sort is stable: equal destinations keep ascending original position.
  SG_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, indices, keys_sorted, (const int32_t *)iota, t_perm, nnz,
                                          0, key_bits(n_nb), st));
  sg::count_launch(4);
  indptr_from_sorted<<<ceil_div(n_nb + 1, 256), 256, 0, st>>>(t_indptr, keys_sorted, nnz, n_nb);
  SG_LAUNCHED("indptr_from_sorted");
  gather_i32<<<grid_for(nnz), 256, 0, st>>>(t_seg, seg_ids, t_perm, nnz);
  SG_LAUNCHED("gather_i32");
  return SG_OK;
}

size_t sg_plan_bytes(int n_seg, int nnz, int chunk) {
  if (n_seg < 0 || nnz < 0 || chunk <= 0) return 0;
  return plan_off_scan(n_seg, nnz, chunk) + 3 * align_up((size_t)(n_seg + 1) * sizeof(int32_t), 64) +
         scan_ws_bytes(n_seg) + 64;
}

size_t sg_plan_partial_rows(int n_seg, int nnz, int chunk) {
  (void)n_seg;
  if (nnz < 0 || chunk <= 0) return 0;
  // a segment is split only if len > chunk, and then into ceil(len/chunk) <= 2*len/chunk rows
  return 2 * ((size_t)nnz / chunk) + 1;
}

int sg_plan_build(void *plan, size_t plan_bytes, const int32_t *indptr, int n_seg, int nnz, int chunk,
                  sg_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_REQUIRE(plan && indptr, "sg_plan_build: null pointer");
  SG_REQUIRE(n_seg >= 0 && nnz >= 0 && chunk > 0, "sg_plan_build: bad sizes (n_seg=%d nnz=%d chunk=%d)", n_seg, nnz, chunk);
  SG_REQUIRE((reinterpret_cast<uintptr_t>(plan) & 15) == 0, "sg_plan_build: plan buffer must be 16-byte aligned");
  const size_t need = sg_plan_bytes(n_seg, nnz, chunk);
  if (plan_bytes < need) return sg::fail(SG_ERR_WORKSPACE, "sg_plan_build: plan buffer %zu < %zu bytes", plan_bytes, need);
  char *base = static_cast<char *>(plan);
  PlanHeader *hdr = reinterpret_cast<PlanHeader *>(base);
  int4 *items = reinterpret_cast<int4 *>(base + plan_off_items());
  int4 *longs = reinterpret_cast<int4 *>(base + plan_off_longs(n_seg, nnz, chunk));
  const size_t arr = align_up((size_t)(n_seg + 1) * sizeof(int32_t), 64);
  char *scan_base = base + plan_off_scan(n_seg, nnz, chunk);
  int32_t *a_item = reinterpret_cast<int32_t *>(scan_base);
  int32_t *a_long = reinterpret_cast<int32_t *>(scan_base + arr);
  int32_t *a_part = reinterpret_cast<int32_t *>(scan_base + 2 * arr);
  void *scan_ws = scan_base + 3 * arr;
  SG_CUDA(cudaMemsetAsync(hdr, 0, sizeof(PlanHeader), st));
  if (n_seg == 0) return SG_OK;
  // the scan region (3 count arrays + scan scratch in the first version) now only holds 3 x n_tiles tile sums
  (void)a_long; (void)a_part; (void)scan_ws;
  const int n_tiles = ceil_div(n_seg, kScanTile);
  int32_t *tile_sums = a_item;
  plan_tile_sums<<<n_tiles, kScanThreads, 0, st>>>(indptr, n_seg, chunk, tile_sums, n_tiles);
  SG_LAUNCHED("plan_tile_sums");
  plan_scan_tiles<<<1, kScanThreads, 0, st>>>(tile_sums, n_tiles, hdr);
  SG_LAUNCHED("plan_scan_tiles");
  plan_fill_fused<<<n_tiles, kScanThreads, 0, st>>>(hdr, items, longs, indptr, n_seg, nnz, chunk, tile_sums, n_tiles,
                                                    (int)plan_cap_items(n_seg, nnz, chunk), (int)plan_cap_long(nnz, chunk));
  SG_LAUNCHED("plan_fill_fused");
  return SG_OK;
}

int sg_multilink_transpose_finish(int32_t *t_src, float *t_w, const int32_t *t_perm, const int32_t *t_seg,
                                  const float *support, int R, int n_dst, int nnz, sg_stream_t stream) {
  SG_REQUIRE(R > 0 && n_dst >= 0 && nnz >= 0, "sg_multilink_transpose_finish: bad sizes");
  if (nnz == 0) return SG_OK;
  SG_REQUIRE(t_src && t_w && t_perm && t_seg && support, "sg_multilink_transpose_finish: null pointer");
  multilink_finish<<<grid_for(nnz), 256, 0, (cudaStream_t)stream>>>(t_src, t_w, t_perm, t_seg, support, R, n_dst, nnz);
  SG_LAUNCHED("multilink_finish");
  return SG_OK;
}

int sg_upload_segments(void *const *dst_device, const void *const *src_pinned_host, const size_t *bytes, int n,
                       sg_stream_t stream) {
  SG_REQUIRE(n >= 0, "sg_upload_segments: negative count");
  SG_REQUIRE(n == 0 || (dst_device && src_pinned_host && bytes), "sg_upload_segments: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  for (int first = 0; first < n; first += kUploadMaxSeg) {
    UploadArgs a;
    a.n = 0;
    long long total = 0;
    for (int k = first; k < n && a.n < kUploadMaxSeg; ++k) {
      if (bytes[k] == 0) continue;
      SG_REQUIRE(dst_device[k] && src_pinned_host[k], "sg_upload_segments: null segment pointer");
      SG_REQUIRE((bytes[k] & 3) == 0 && (reinterpret_cast<uintptr_t>(dst_device[k]) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(src_pinned_host[k]) & 3) == 0,
                 "sg_upload_segments: segments must be 4-byte aligned multiples of 4 bytes");
      a.words[a.n] = (long long)(bytes[k] / 4);
      total += (a.words[a.n] + 3) / 4;
      a.src[a.n] = static_cast<const uint32_t *>(src_pinned_host[k]);
      a.dst[a.n] = static_cast<uint32_t *>(dst_device[k]);
      a.end[a.n] = total;
      ++a.n;
    }
    if (a.n == 0) continue;
    long long blocks = ceil_div<long long>(total, 256);   // one 16-byte unit per thread: every request in flight at once
    const long long cap = (long long)num_sms() * 4;      // enough requests in flight for PCIe, few SMs taken
    if (blocks > cap) blocks = cap;
    upload_segments_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    SG_LAUNCHED("upload_segments_kernel");
  }
  return SG_OK;
}

size_t sg_unique_inverse_ws_bytes(int n) {
  if (n <= 0) return 64;
  size_t cub_bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                                  (const int32_t *)nullptr, (int32_t *)nullptr, n, 0, 32);
  if (e != cudaSuccess) {
    sg::fail(SG_ERR_CUDA, "sg_unique_inverse_ws_bytes: CUB size query failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return 0;
  }
  return 8 * align_up((size_t)n * sizeof(int32_t), 256) + align_up(cub_bytes, 256) + scan_ws_bytes(n) + 512;
}

int sg_unique_inverse(int32_t *uniq, int32_t *inverse, int32_t *n_unique, const int32_t *data, int n, void *ws,
                      size_t ws_bytes, sg_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_REQUIRE(n >= 0, "sg_unique_inverse: negative size");
  SG_REQUIRE(n_unique, "sg_unique_inverse: null n_unique");
  if (n == 0) { SG_CUDA(cudaMemsetAsync(n_unique, 0, sizeof(int32_t), st)); return SG_OK; }
  SG_REQUIRE(uniq && inverse && data && ws, "sg_unique_inverse: null pointer");
  const size_t need = sg_unique_inverse_ws_bytes(n);
  if (need == 0) return SG_ERR_CUDA;
  if (ws_bytes < need) return sg::fail(SG_ERR_WORKSPACE, "sg_unique_inverse: workspace %zu < %zu bytes", ws_bytes, need);
  const size_t stride = align_up((size_t)n * sizeof(int32_t), 256);
  char *base = static_cast<char *>(ws);
  int32_t *iota = reinterpret_cast<int32_t *>(base);
  int32_t *ks = reinterpret_cast<int32_t *>(base + stride);
  int32_t *pos = reinterpret_cast<int32_t *>(base + 2 * stride);
  int32_t *head = reinterpret_cast<int32_t *>(base + 3 * stride);
  int32_t *first_pos = reinterpret_cast<int32_t *>(base + 4 * stride);
  int32_t *run_iota = reinterpret_cast<int32_t *>(base + 5 * stride);
  int32_t *fp_sorted = reinterpret_cast<int32_t *>(base + 6 * stride);
  int32_t *order = reinterpret_cast<int32_t *>(base + 7 * stride);
  void *cub_ws = base + 8 * stride;
  size_t cub_bytes = 0;
  SG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                          (const int32_t *)nullptr, (int32_t *)nullptr, n, 0, 32));
  void *scan_ws = static_cast<char *>(cub_ws) + align_up(cub_bytes, 256);
  int32_t *rank = head;  // head[] is dead once its scan (in place, exclusive) has been consumed... keep separate: reuse iota
  iota_kernel<<<grid_for(n), 256, 0, st>>>(iota, n);
  SG_LAUNCHED("iota_kernel");
  SG_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, data, ks, (const int32_t *)iota, pos, n, 0, 32, st));
  sg::count_launch(4);
  run_heads_kernel<<<grid_for(n), 256, 0, st>>>(head, ks, n);
  SG_LAUNCHED("run_heads_kernel");
  int rc = exclusive_scan_i32(head, head, n, n_unique, scan_ws, st);  // head[j] := number of heads before j
  if (rc != SG_OK) return rc;
  // sentinel first positions (INT_MAX) for the unused tail so that the second sort keeps real runs in front
  SG_CUDA(cudaMemsetAsync(first_pos, 0x7f, stride, st));
  SG_CUDA(cudaMemsetAsync(run_iota, 0, stride, st));
  run_first_pos_kernel<<<grid_for(n), 256, 0, st>>>(first_pos, run_iota, head, ks, pos, n);
  SG_LAUNCHED("run_first_pos_kernel");
  SG_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, (const int32_t *)first_pos, fp_sorted,
                                          (const int32_t *)run_iota, order, n, 0, 32, st));
  sg::count_launch(4);
  rank = iota;  // iota[] is no longer needed
  rank_scatter_kernel<<<grid_for(n), 256, 0, st>>>(rank, order, n_unique);
  SG_LAUNCHED("rank_scatter_kernel");
  unique_write_kernel<<<grid_for(n), 256, 0, st>>>(uniq, inverse, rank, head, ks, pos, n);
  SG_LAUNCHED("unique_write_kernel");
  return SG_OK;
}

}  // extern "C"
