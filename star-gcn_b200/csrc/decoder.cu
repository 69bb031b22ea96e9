// Masked-embedding reconstruction decoder and rating head: the small kernels around the Dense
// layers (which run on the tcgen05 GEMM, gemm.cu).  Reference: experiments/STAR-GCN.py
//   get_embed :264-300   embed_maps :226-246 / :441-454   recon loss :618-628   rating head :428-438
//   InnerProductLayer mxgraph/layers/layers.py:217-222     L2Loss :611-616
// Every reduction here is two-stage with a fixed order: results are bit-identical run to run.
#include "common.cuh"

namespace sg {

constexpr int kRedBlocks = 296;   // 2 x 148 SMs
constexpr int kRedThreads = 256;

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float warp_part[kRedThreads / 32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warp_part[wid] = v;
  __syncthreads();
  float t = 0.f;
  if (wid == 0) {
    t = lane < kRedThreads / 32 ? warp_part[lane] : 0.f;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

// D1: one group of lanes per output row
__global__ void __launch_bounds__(256) masked_embed_fwd_kernel(float *__restrict__ out, int32_t *__restrict__ eff_ids,
                                                               const float *__restrict__ table,
                                                               const int32_t *__restrict__ ids,
                                                               const int32_t *__restrict__ noise, int n, int n_table, int D) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const bool v4 = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(table)) & 15) == 0;
  for (int i = warp; i < n; i += n_warps) {
    int id = __ldg(ids + i);
    if (noise) id = (id >= 0 && id < n_table) ? __ldg(noise + id) : -1;
    if (id >= n_table) id = n_table - 1;  // mx.nd.take / Embedding clip mode
    if (lane == 0 && eff_ids) eff_ids[i] = id;
    float *o = out + (long long)i * D;
    if (v4) {
      const float4 *src = reinterpret_cast<const float4 *>(table + (long long)(id < 0 ? 0 : id) * D);
      for (int c = lane; c < D / 4; c += 32)
        reinterpret_cast<float4 *>(o)[c] = id < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(src + c);
    } else {
      const float *src = table + (long long)(id < 0 ? 0 : id) * D;
      for (int c = lane; c < D; c += 32) o[c] = id < 0 ? 0.f : __ldg(src + c);
    }
  }
}

// D3 stage 1: per-block partial of scale * sum (a-b)^2 over a fixed slice
__global__ void __launch_bounds__(kRedThreads) sq_err_partial_kernel(float *__restrict__ partial, const float *__restrict__ a,
                                                                     const float *__restrict__ b, long long n_elem) {
  const long long per_block = (n_elem + gridDim.x - 1) / gridDim.x;
  const long long lo = per_block * blockIdx.x, hi = min(lo + per_block, n_elem);
  float acc = 0.f;
  for (long long t = lo + threadIdx.x; t < hi; t += kRedThreads) {
    const float d = __ldg(a + t) - __ldg(b + t);
    acc = fmaf(d, d, acc);
  }
  const float s = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kRedThreads) final_sum_kernel(float *__restrict__ out, const float *__restrict__ partial,
                                                                int n_partial, float scale) {
  float acc = 0.f;
  for (int t = threadIdx.x; t < n_partial; t += kRedThreads) acc += partial[t];
  const float s = block_sum(acc);
  if (threadIdx.x == 0) out[0] = s * scale;
}

__global__ void __launch_bounds__(256) sq_err_bwd_kernel(float *__restrict__ ga, float *__restrict__ gb,
                                                         const float *__restrict__ a, const float *__restrict__ b,
                                                         const float *__restrict__ gloss, long long n_elem, float scale) {
  const float k = 2.f * scale * __ldg(gloss);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_elem; t += (long long)gridDim.x * blockDim.x) {
    const float g = k * (__ldg(a + t) - __ldg(b + t));
    if (ga) ga[t] = g;
    if (gb) gb[t] = -g;
  }
}

// D4: one warp per row
__global__ void __launch_bounds__(256) rowdot_fwd_kernel(float *__restrict__ out, const float *__restrict__ a,
                                                         const float *__restrict__ b, int n, int D) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n; i += n_warps) {
    const float *pa = a + (long long)i * D, *pb = b + (long long)i * D;
    float acc = 0.f;
    for (int c = lane; c < D; c += 32) acc = fmaf(__ldg(pa + c), __ldg(pb + c), acc);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) out[i] = acc;
  }
}

__global__ void __launch_bounds__(256) rowdot_bwd_kernel(float *__restrict__ ga, float *__restrict__ gb,
                                                         const float *__restrict__ gout, const float *__restrict__ a,
                                                         const float *__restrict__ b, int n, int D) {
  const long long total = (long long)n * D;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const float g = __ldg(gout + t / D);
    if (ga) ga[t] = g * __ldg(b + t);
    if (gb) gb[t] = g * __ldg(a + t);
  }
}

// column sums, stage 1: block (x: column chunk of 32, y: row slice) -> partial[y][col]
__global__ void __launch_bounds__(256) colsum_partial_kernel(float *__restrict__ partial, const float *__restrict__ x_hi,
                                                             const float *__restrict__ x_lo, int M, int N, int ld) {
  __shared__ float tile[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;  // 32 columns x 8 row lanes
  const int col = blockIdx.x * 32 + cx;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(r0 + rows_per, M);
  float acc = 0.f;
  if (col < N) {
    for (int r = r0 + ry; r < r1; r += 8) {
      float v = __ldg(x_hi + (long long)r * ld + col);
      if (x_lo) v += __ldg(x_lo + (long long)r * ld + col);
      acc += v;
    }
  }
  tile[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && col < N) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += tile[j][cx];
    partial[(long long)blockIdx.y * N + col] = s;
  }
}

__global__ void __launch_bounds__(256) colsum_final_kernel(float *__restrict__ out, const float *__restrict__ partial,
                                                           int N, int slices) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int j = 0; j < slices; ++j) s += partial[(long long)j * N + c];
  out[c] = s;
}

constexpr int kColsumSlices = 64;

static inline int grid_ew2(long long n, int per_block = 256) {
  long long gsz = ceil_div<long long>(n > 0 ? n : 1, per_block);
  long long cap = (long long)num_sms() * 16;
  return (int)(gsz < cap ? gsz : cap);
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_masked_embed_fwd(float *out, int32_t *eff_ids, const float *table, const int32_t *ids, const int32_t *noise,
                        int n, int n_table, int D, sg_stream_t stream) {
  SG_REQUIRE(n >= 0 && n_table > 0 && D > 0, "sg_masked_embed_fwd: bad sizes (n=%d n_table=%d D=%d)", n, n_table, D);
  if (n == 0) return SG_OK;
  SG_REQUIRE(out && table && ids, "sg_masked_embed_fwd: null pointer");
  masked_embed_fwd_kernel<<<grid_ew2(n, 8), 256, 0, (cudaStream_t)stream>>>(out, eff_ids, table, ids, noise, n, n_table, D);
  SG_LAUNCHED("masked_embed_fwd_kernel");
  return SG_OK;
}

size_t sg_reduce_ws_bytes(void) { return kRedBlocks * sizeof(float); }

int sg_sq_err_fwd(float *loss, const float *a, const float *b, long long n_elem, float scale, void *ws, sg_stream_t stream) {
  SG_REQUIRE(n_elem >= 0, "sg_sq_err_fwd: negative size");
  SG_REQUIRE(loss && ws && (n_elem == 0 || (a && b)), "sg_sq_err_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  sq_err_partial_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(static_cast<float *>(ws), a, b, n_elem);
  SG_LAUNCHED("sq_err_partial_kernel");
  final_sum_kernel<<<1, kRedThreads, 0, st>>>(loss, static_cast<const float *>(ws), kRedBlocks, scale);
  SG_LAUNCHED("final_sum_kernel");
  return SG_OK;
}

int sg_sq_err_bwd(float *ga, float *gb, const float *a, const float *b, const float *gloss, long long n_elem, float scale,
                  sg_stream_t stream) {
  SG_REQUIRE(n_elem >= 0, "sg_sq_err_bwd: negative size");
  if (n_elem == 0 || (!ga && !gb)) return SG_OK;
  SG_REQUIRE(a && b && gloss, "sg_sq_err_bwd: null pointer");
  sq_err_bwd_kernel<<<grid_ew2(n_elem), 256, 0, (cudaStream_t)stream>>>(ga, gb, a, b, gloss, n_elem, scale);
  SG_LAUNCHED("sq_err_bwd_kernel");
  return SG_OK;
}

int sg_rowdot_fwd(float *out, const float *a, const float *b, int n, int D, sg_stream_t stream) {
  SG_REQUIRE(n >= 0 && D > 0, "sg_rowdot_fwd: bad sizes");
  if (n == 0) return SG_OK;
  SG_REQUIRE(out && a && b, "sg_rowdot_fwd: null pointer");
  rowdot_fwd_kernel<<<grid_ew2(n, 8), 256, 0, (cudaStream_t)stream>>>(out, a, b, n, D);
  SG_LAUNCHED("rowdot_fwd_kernel");
  return SG_OK;
}

int sg_rowdot_bwd(float *ga, float *gb, const float *gout, const float *a, const float *b, int n, int D, sg_stream_t stream) {
  SG_REQUIRE(n >= 0 && D > 0, "sg_rowdot_bwd: bad sizes");
  if (n == 0 || (!ga && !gb)) return SG_OK;
  SG_REQUIRE(gout && a && b, "sg_rowdot_bwd: null pointer");
  rowdot_bwd_kernel<<<grid_ew2((long long)n * D), 256, 0, (cudaStream_t)stream>>>(ga, gb, gout, a, b, n, D);
  SG_LAUNCHED("rowdot_bwd_kernel");
  return SG_OK;
}

size_t sg_colsum_ws_bytes(int N) { return N > 0 ? (size_t)kColsumSlices * (size_t)N * sizeof(float) : 0; }

int sg_colsum(float *out, const float *x_hi, const float *x_lo, int M, int N, int ld, void *ws, sg_stream_t stream) {
  SG_REQUIRE(M >= 0 && N > 0 && ld >= N, "sg_colsum: bad sizes");
  SG_REQUIRE(out && ws && (M == 0 || x_hi), "sg_colsum: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div(N, 32), (unsigned)kColsumSlices);
  colsum_partial_kernel<<<grid, 256, 0, st>>>(static_cast<float *>(ws), x_hi, x_lo, M, N, ld);
  SG_LAUNCHED("colsum_partial_kernel");
  colsum_final_kernel<<<ceil_div(N, 256), 256, 0, st>>>(out, static_cast<const float *>(ws), N, kColsumSlices);
  SG_LAUNCHED("colsum_final_kernel");
  return SG_OK;
}

}  // extern "C"
