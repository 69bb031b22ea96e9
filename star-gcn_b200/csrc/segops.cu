// C-ABI entry points of the segment operators (see include/stargcn_b200.h) plus the small
// non-gather kernels: contiguous segment reduce / broadcast / softmax, take_k_corr (SDDMM),
// and seg_pool max with arg-max.  The gather-shaped operators route to gather.cu.
#include <cfloat>

#include "common.cuh"
#include "gather.cuh"

namespace sg {

__device__ __forceinline__ int seg_of_pos2(const int32_t *__restrict__ indptr, int n_seg, int p) {
  int lo = 0, hi = n_seg;  // first s with indptr[s+1] > p
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(indptr + mid + 1) > p) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

// ---- A6: one warp per (batch, segment), lanes stride the contiguous run ----
__global__ void __launch_bounds__(256) seg_reduce_kernel(float *__restrict__ dst, const float *__restrict__ data,
                                                         const int32_t *__restrict__ indptr, int B, int nnz, int n_seg,
                                                         int type, int req) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = warp; t < (long long)B * n_seg; t += n_warps) {
    const int b = (int)(t / n_seg), s = (int)(t - (long long)b * n_seg);
    const int lo = __ldg(indptr + s), hi = __ldg(indptr + s + 1);
    const float *row = data + (long long)b * nnz;
    float acc = type == SG_REDUCE_MAX ? -FLT_MAX : (type == SG_REDUCE_MIN ? FLT_MAX : 0.f);
    for (int p = lo + lane; p < hi; p += 32) {
      float v = __ldg(row + p);
      acc = type == SG_REDUCE_SUM ? acc + v : (type == SG_REDUCE_MAX ? fmaxf(acc, v) : fminf(acc, v));
    }
    acc = type == SG_REDUCE_SUM ? warp_sum(acc) : (type == SG_REDUCE_MAX ? warp_max(acc) : warp_min(acc));
    if (lane == 0) {
      float *o = dst + (long long)b * n_seg + s;
      *o = req == SG_REQ_ADD ? *o + acc : acc;
    }
  }
}

__global__ void __launch_bounds__(256) seg_broadcast_kernel(float *__restrict__ dst, const float *__restrict__ lhs,
                                                            const float *__restrict__ rhs,
                                                            const int32_t *__restrict__ indptr, int B, int nnz,
                                                            int n_seg, int op, int req) {
  const long long total = (long long)B * nnz;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(t / nnz), p = (int)(t - (long long)b * nnz);
    const int s = seg_of_pos2(indptr, n_seg, p);
    const bool covered = s < n_seg && __ldg(indptr + s) <= p;
    if (!covered) {  // outside every segment: 0 under WRITE (reference memset), untouched under ADD
      if (req != SG_REQ_ADD) dst[t] = 0.f;
      continue;
    }
    const float r = __ldg(rhs + (long long)b * n_seg + s);
    const float l = lhs ? __ldg(lhs + t) : 0.f;
    float v;
    switch (op) {
      case SG_BCAST_ADD: v = l + r; break;
      case SG_BCAST_MUL: v = l * r; break;
      case SG_BCAST_TO: v = r; break;
      case SG_BCAST_SUB: v = l - r; break;
      default: v = l / r; break;
    }
    dst[t] = req == SG_REQ_ADD ? dst[t] + v : v;
  }
}

__global__ void __launch_bounds__(256) seg_softmax_fwd_kernel(float *__restrict__ dst, const float *__restrict__ data,
                                                              const int32_t *__restrict__ indptr, int B, int nnz,
                                                              int n_seg) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = warp; t < (long long)B * n_seg; t += n_warps) {
    const int b = (int)(t / n_seg), s = (int)(t - (long long)b * n_seg);
    const int lo = __ldg(indptr + s), hi = __ldg(indptr + s + 1);
    const float *x = data + (long long)b * nnz;
    float *y = dst + (long long)b * nnz;
    float mx = -FLT_MAX;
    for (int p = lo + lane; p < hi; p += 32) mx = fmaxf(mx, __ldg(x + p));
    mx = warp_max(mx);
    float sum = 0.f;
    for (int p = lo + lane; p < hi; p += 32) sum += expf(__ldg(x + p) - mx);
    sum = warp_sum(sum);
    for (int p = lo + lane; p < hi; p += 32) y[p] = expf(__ldg(x + p) - mx) / sum;
  }
}

__global__ void __launch_bounds__(256) seg_softmax_bwd_kernel(float *__restrict__ dst, const float *__restrict__ og,
                                                              const float *__restrict__ val,
                                                              const int32_t *__restrict__ indptr, int B, int nnz,
                                                              int n_seg, int req) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = warp; t < (long long)B * n_seg; t += n_warps) {
    const int b = (int)(t / n_seg), s = (int)(t - (long long)b * n_seg);
    const int lo = __ldg(indptr + s), hi = __ldg(indptr + s + 1);
    const long long off = (long long)b * nnz;
    float dot = 0.f;
    for (int p = lo + lane; p < hi; p += 32) dot += __ldg(og + off + p) * __ldg(val + off + p);
    dot = warp_sum(dot);
    for (int p = lo + lane; p < hi; p += 32) {
      float g = __ldg(val + off + p) * (__ldg(og + off + p) - dot);
      dst[off + p] = req == SG_REQ_ADD ? dst[off + p] + g : g;
    }
  }
}

// ---- A4: SDDMM, one LPR-lane group per edge ----
template <int LPR>
__global__ void __launch_bounds__(256) take_k_corr_kernel(float *__restrict__ dst, const float *__restrict__ e1,
                                                          const float *__restrict__ e2,
                                                          const int32_t *__restrict__ idx,
                                                          const int32_t *__restrict__ indptr, int K, int n_node,
                                                          int n_nb, int nnz, int F, int req) {
  const int lane = threadIdx.x & (LPR - 1);
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const long long n_groups = ((long long)gridDim.x * blockDim.x) / LPR;
  const long long total = (long long)K * nnz;
  // every lane of a warp must reach the shuffles: iterate on a warp-uniform bound
  const long long rounds = (total + n_groups - 1) / n_groups;
  for (long long rnd = 0; rnd < rounds; ++rnd) {
    const long long t = rnd * n_groups + group;
    float acc = 0.f;
    int k = 0, p = 0;
    const bool live = t < total;
    if (live) {
      k = (int)(t / nnz); p = (int)(t - (long long)k * nnz);
      const int s = seg_of_pos2(indptr, n_node, p);
      const float *a = e1 + ((long long)k * n_node + s) * F;
      const float *b = e2 + ((long long)k * n_nb + __ldg(idx + p)) * F;
      for (int c = lane; c < F; c += LPR) acc = fmaf(__ldg(a + c), __ldg(b + c), acc);
    }
#pragma unroll
    for (int d = LPR / 2; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (live && lane == 0) {
      float *o = dst + (long long)k * nnz + p;
      *o = req == SG_REQ_ADD ? *o + acc : acc;
    }
  }
}

// ---- A5 max: one thread per (batch, segment, channel); strict '>' keeps the first position ----
__global__ void __launch_bounds__(256) seg_pool_max_fwd_kernel(float *__restrict__ dst, int32_t *__restrict__ argmax,
                                                               const float *__restrict__ data,
                                                               const int32_t *__restrict__ idx,
                                                               const int32_t *__restrict__ indptr, int B, int n_seg,
                                                               int n_nb, int F) {
  const long long total = (long long)B * n_seg * F;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % F);
    const long long bs = t / F;
    const int s = (int)(bs % n_seg), b = (int)(bs / n_seg);
    const int lo = __ldg(indptr + s), hi = __ldg(indptr + s + 1);
    float best = hi > lo ? -FLT_MAX : 0.f;
    int arg = -1;
    const float *src = data + (long long)b * n_nb * F + c;
    for (int p = lo; p < hi; ++p) {
      float v = __ldg(src + (long long)__ldg(idx + p) * F);
      if (v > best) { best = v; arg = p; }
    }
    dst[t] = best;
    if (argmax) argmax[t] = arg;
  }
}

__global__ void __launch_bounds__(256) seg_pool_max_bwd_kernel(float *__restrict__ gdata, const float *__restrict__ gout,
                                                               const int32_t *__restrict__ argmax,
                                                               const int32_t *__restrict__ t_indptr,
                                                               const int32_t *__restrict__ t_perm,
                                                               const int32_t *__restrict__ t_seg, int B, int n_seg,
                                                               int n_nb, int F, int req) {
  const long long total = (long long)B * n_nb * F;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % F);
    const long long bn = t / F;
    const int n = (int)(bn % n_nb), b = (int)(bn / n_nb);
    float acc = 0.f;
    for (int q = __ldg(t_indptr + n); q < __ldg(t_indptr + n + 1); ++q) {
      const long long o = ((long long)b * n_seg + __ldg(t_seg + q)) * F + c;
      if (__ldg(argmax + o) == __ldg(t_perm + q)) acc += __ldg(gout + o);
    }
    gdata[t] = req == SG_REQ_ADD ? gdata[t] + acc : acc;
  }
}

static inline int grid_1d(long long work_items, int per_block) {
  long long g = ceil_div<long long>(work_items > 0 ? work_items : 1, per_block);
  long long cap = (long long)num_sms() * 32;
  return (int)(g < cap ? g : cap);
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_weighted_pool_fwd(float *dst, const float *data, const float *weights, const int32_t *indices,
                         const int32_t *indptr, int K, int n_seg, int n_nb, int nnz, int F, int req,
                         const void *plan, int plan_chunk, float *partial, sg_stream_t stream) {
  SG_REQUIRE(valid_req(req), "sg_weighted_pool_fwd: bad req %d", req);
  if (req == SG_REQ_NULL) return SG_OK;
  SG_REQUIRE(K >= 0 && n_seg >= 0 && n_nb >= 0 && nnz >= 0 && F >= 0, "sg_weighted_pool_fwd: negative size");
  SG_REQUIRE((long long)K * n_nb * F < (1LL << 40), "sg_weighted_pool_fwd: tensor too large");
  if (K == 0 || n_seg == 0 || F == 0) return SG_OK;
  SG_REQUIRE(dst && indptr && (nnz == 0 || (data && weights && indices)), "sg_weighted_pool_fwd: null pointer");
  GatherArgs a;
  a.out = dst; a.out_batch_stride = (long long)n_seg * F; a.ld_out = F; a.n_out_rows = n_seg;
  a.src = data; a.src_batch_stride = (long long)n_nb * F; a.ld_src = F;
  a.w = weights; a.w_batch_stride = nnz;
  a.idx = indices; a.indptr = indptr; a.F = F; a.req = req;
  a.plan_chunk = plan_chunk; a.partial = partial;
  a.partial_batch_stride = plan ? (long long)sg_plan_partial_rows(n_seg, nnz, plan_chunk) * F : 0;
  return run_gather(a, K, n_seg, nnz, plan, (cudaStream_t)stream);
}

int sg_weighted_pool_bwd_data(float *gdata, const float *gout, const float *weights, const int32_t *t_indptr,
                              const int32_t *t_perm, const int32_t *t_seg, int K, int n_seg, int n_nb, int nnz,
                              int F, int req, const void *t_plan, int plan_chunk, float *partial,
                              sg_stream_t stream) {
  SG_REQUIRE(valid_req(req), "sg_weighted_pool_bwd_data: bad req %d", req);
  if (req == SG_REQ_NULL) return SG_OK;
  SG_REQUIRE(K >= 0 && n_seg >= 0 && n_nb >= 0 && nnz >= 0 && F >= 0, "sg_weighted_pool_bwd_data: negative size");
  if (K == 0 || n_nb == 0 || F == 0) return SG_OK;
  SG_REQUIRE(gdata && t_indptr && (nnz == 0 || (gout && weights && t_perm && t_seg)),
             "sg_weighted_pool_bwd_data: null pointer");
  GatherArgs a;
  a.out = gdata; a.out_batch_stride = (long long)n_nb * F; a.ld_out = F; a.n_out_rows = n_nb;
  a.src = gout; a.src_batch_stride = (long long)n_seg * F; a.ld_src = F;
  a.w = weights; a.w_batch_stride = nnz; a.perm = t_perm;
  a.idx = t_seg; a.indptr = t_indptr; a.F = F; a.req = req;
  a.plan_chunk = plan_chunk; a.partial = partial;
  a.partial_batch_stride = t_plan ? (long long)sg_plan_partial_rows(n_nb, nnz, plan_chunk) * F : 0;
  return run_gather(a, K, n_nb, nnz, t_plan, (cudaStream_t)stream);
}

int sg_take_k_corr(float *dst, const float *embed1, const float *embed2, const int32_t *indices,
                   const int32_t *indptr, int K, int n_node, int n_nb, int nnz, int F, int req,
                   sg_stream_t stream) {
  SG_REQUIRE(valid_req(req), "sg_take_k_corr: bad req %d", req);
  if (req == SG_REQ_NULL) return SG_OK;
  SG_REQUIRE(K >= 0 && n_node >= 0 && n_nb >= 0 && nnz >= 0 && F >= 0, "sg_take_k_corr: negative size");
  if (K == 0 || nnz == 0) return SG_OK;
  SG_REQUIRE(dst && embed1 && embed2 && indices && indptr, "sg_take_k_corr: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)K * nnz;
  if (F <= 8) {
    take_k_corr_kernel<4><<<grid_1d(total, 256 / 4), 256, 0, st>>>(dst, embed1, embed2, indices, indptr, K, n_node, n_nb, nnz, F, req);
  } else if (F <= 64) {
    take_k_corr_kernel<16><<<grid_1d(total, 256 / 16), 256, 0, st>>>(dst, embed1, embed2, indices, indptr, K, n_node, n_nb, nnz, F, req);
  } else {
    take_k_corr_kernel<32><<<grid_1d(total, 256 / 32), 256, 0, st>>>(dst, embed1, embed2, indices, indptr, K, n_node, n_nb, nnz, F, req);
  }
  SG_LAUNCHED("take_k_corr_kernel");
  return SG_OK;
}

int sg_seg_pool_fwd(float *dst, int32_t *argmax, const float *data, const int32_t *indices, const int32_t *indptr,
                    int B, int n_seg, int n_nb, int nnz, int F, int pool_type, const void *plan, int plan_chunk,
                    float *partial, sg_stream_t stream) {
  SG_REQUIRE(pool_type == SG_POOL_SUM || pool_type == SG_POOL_MEAN || pool_type == SG_POOL_MAX,
             "sg_seg_pool_fwd: bad pool_type %d", pool_type);
  SG_REQUIRE(B >= 0 && n_seg >= 0 && n_nb >= 0 && nnz >= 0 && F >= 0, "sg_seg_pool_fwd: negative size");
  if (B == 0 || n_seg == 0 || F == 0) return SG_OK;
  SG_REQUIRE(dst && indptr && (nnz == 0 || (data && indices)), "sg_seg_pool_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (pool_type == SG_POOL_MAX) {
    const long long total = (long long)B * n_seg * F;
    seg_pool_max_fwd_kernel<<<grid_1d(total, 256), 256, 0, st>>>(dst, argmax, data, indices, indptr, B, n_seg, n_nb, F);
    SG_LAUNCHED("seg_pool_max_fwd_kernel");
    return SG_OK;
  }
  GatherArgs a;
  a.out = dst; a.out_batch_stride = (long long)n_seg * F; a.ld_out = F; a.n_out_rows = n_seg;
  a.src = data; a.src_batch_stride = (long long)n_nb * F; a.ld_src = F;
  a.idx = indices; a.indptr = indptr; a.F = F; a.req = SG_REQ_WRITE;
  a.mean = pool_type == SG_POOL_MEAN;
  a.plan_chunk = plan_chunk; a.partial = partial;
  a.partial_batch_stride = plan ? (long long)sg_plan_partial_rows(n_seg, nnz, plan_chunk) * F : 0;
  return run_gather(a, B, n_seg, nnz, plan, st);
}

int sg_seg_pool_bwd(float *gdata, const float *gout, const int32_t *argmax, const int32_t *indptr,
                    const int32_t *t_indptr, const int32_t *t_perm, const int32_t *t_seg, int B, int n_seg, int n_nb,
                    int nnz, int F, int pool_type, int req, const void *t_plan, int plan_chunk, float *partial,
                    sg_stream_t stream) {
  SG_REQUIRE(valid_req(req), "sg_seg_pool_bwd: bad req %d", req);
  SG_REQUIRE(pool_type == SG_POOL_SUM || pool_type == SG_POOL_MEAN || pool_type == SG_POOL_MAX,
             "sg_seg_pool_bwd: bad pool_type %d", pool_type);
  if (req == SG_REQ_NULL) return SG_OK;
  SG_REQUIRE(B >= 0 && n_seg >= 0 && n_nb >= 0 && nnz >= 0 && F >= 0, "sg_seg_pool_bwd: negative size");
  if (B == 0 || n_nb == 0 || F == 0) return SG_OK;
  SG_REQUIRE(gdata && t_indptr && (nnz == 0 || (gout && t_perm && t_seg)), "sg_seg_pool_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (pool_type == SG_POOL_MAX) {
    SG_REQUIRE(argmax || nnz == 0, "sg_seg_pool_bwd: max pooling needs argmax");
    const long long total = (long long)B * n_nb * F;
    seg_pool_max_bwd_kernel<<<grid_1d(total, 256), 256, 0, st>>>(gdata, gout, argmax, t_indptr, t_perm, t_seg, B, n_seg, n_nb, F, req);
    SG_LAUNCHED("seg_pool_max_bwd_kernel");
    return SG_OK;
  }
  GatherArgs a;
  a.out = gdata; a.out_batch_stride = (long long)n_nb * F; a.ld_out = F; a.n_out_rows = n_nb;
  a.src = gout; a.src_batch_stride = (long long)n_seg * F; a.ld_src = F;
  a.idx = t_seg; a.indptr = t_indptr; a.F = F; a.req = req;
  if (pool_type == SG_POOL_MEAN) {
    SG_REQUIRE(indptr, "sg_seg_pool_bwd: mean pooling needs the forward indptr");
    a.inv_len_indptr = indptr;
  }
  a.plan_chunk = plan_chunk; a.partial = partial;
  a.partial_batch_stride = t_plan ? (long long)sg_plan_partial_rows(n_nb, nnz, plan_chunk) * F : 0;
  return run_gather(a, B, n_nb, nnz, t_plan, st);
}

int sg_seg_reduce(float *dst, const float *data, const int32_t *indptr, int B, int nnz, int n_seg, int reduce_type,
                  int req, sg_stream_t stream) {
  SG_REQUIRE(valid_req(req), "sg_seg_reduce: bad req %d", req);
  SG_REQUIRE(reduce_type == SG_REDUCE_SUM || reduce_type == SG_REDUCE_MAX || reduce_type == SG_REDUCE_MIN,
             "sg_seg_reduce: bad reduce_type %d", reduce_type);
  if (req == SG_REQ_NULL) return SG_OK;
  SG_REQUIRE(B >= 0 && nnz >= 0 && n_seg >= 0, "sg_seg_reduce: negative size");
  if (B == 0 || n_seg == 0) return SG_OK;
  SG_REQUIRE(dst && indptr && (nnz == 0 || data), "sg_seg_reduce: null pointer");
  seg_reduce_kernel<<<grid_1d((long long)B * n_seg, 8), 256, 0, (cudaStream_t)stream>>>(dst, data, indptr, B, nnz, n_seg, reduce_type, req);
  SG_LAUNCHED("seg_reduce_kernel");
  return SG_OK;
}

int sg_seg_broadcast_binary(float *dst, const float *lhs, const float *rhs, const int32_t *indptr, int B, int nnz,
                            int n_seg, int op, int req, sg_stream_t stream) {
  SG_REQUIRE(valid_req(req), "sg_seg_broadcast_binary: bad req %d", req);
  SG_REQUIRE(op >= SG_BCAST_ADD && op <= SG_BCAST_DIV, "sg_seg_broadcast_binary: bad op %d", op);
  if (req == SG_REQ_NULL) return SG_OK;
  SG_REQUIRE(B >= 0 && nnz >= 0 && n_seg >= 0, "sg_seg_broadcast_binary: negative size");
  if (B == 0 || nnz == 0) return SG_OK;
  SG_REQUIRE(dst && rhs && indptr && (lhs || op == SG_BCAST_TO), "sg_seg_broadcast_binary: null pointer");
  seg_broadcast_kernel<<<grid_1d((long long)B * nnz, 256), 256, 0, (cudaStream_t)stream>>>(dst, lhs, rhs, indptr, B, nnz, n_seg, op, req);
  SG_LAUNCHED("seg_broadcast_kernel");
  return SG_OK;
}

int sg_seg_softmax_fwd(float *dst, const float *data, const int32_t *indptr, int B, int nnz, int n_seg,
                       sg_stream_t stream) {
  SG_REQUIRE(B >= 0 && nnz >= 0 && n_seg >= 0, "sg_seg_softmax_fwd: negative size");
  if (B == 0 || nnz == 0) return SG_OK;
  SG_REQUIRE(dst && data && indptr, "sg_seg_softmax_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  SG_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)B * nnz, st));
  if (n_seg == 0) return SG_OK;
  seg_softmax_fwd_kernel<<<grid_1d((long long)B * n_seg, 8), 256, 0, st>>>(dst, data, indptr, B, nnz, n_seg);
  SG_LAUNCHED("seg_softmax_fwd_kernel");
  return SG_OK;
}

int sg_seg_softmax_bwd(float *dst, const float *ograd, const float *val, const int32_t *indptr, int B, int nnz,
                       int n_seg, int req, sg_stream_t stream) {
  SG_REQUIRE(valid_req(req), "sg_seg_softmax_bwd: bad req %d", req);
  if (req == SG_REQ_NULL) return SG_OK;
  SG_REQUIRE(B >= 0 && nnz >= 0 && n_seg >= 0, "sg_seg_softmax_bwd: negative size");
  if (B == 0 || nnz == 0 || n_seg == 0) return SG_OK;
  SG_REQUIRE(dst && ograd && val && indptr, "sg_seg_softmax_bwd: null pointer");
  seg_softmax_bwd_kernel<<<grid_1d((long long)B * n_seg, 8), 256, 0, (cudaStream_t)stream>>>(dst, ograd, val, indptr, B, nnz, n_seg, req);
  SG_LAUNCHED("seg_softmax_bwd_kernel");
  return SG_OK;
}

static int multilink_agg_fwd_impl(float *agg, float *agg_lo, int ld_agg, float *wsum, float *wsum_lo, int wsum_ld,
                                  const float *x, const float *support, const int32_t *end_points,
                                  const int32_t *cat_indptr, int R, int n_dst, int n_nb, int nnz, int D,
                                  const void *plan, int plan_chunk, float *partial, sg_stream_t stream) {
  SG_REQUIRE(R > 0 && n_dst >= 0 && n_nb >= 0 && nnz >= 0 && D > 0, "sg_multilink_agg_fwd: bad sizes");
  SG_REQUIRE((long long)R * n_dst < (1LL << 31), "sg_multilink_agg_fwd: R*n_dst overflows int32");
  if (n_dst == 0) return SG_OK;
  SG_REQUIRE(agg && cat_indptr && (nnz == 0 || (x && support && end_points)), "sg_multilink_agg_fwd: null pointer");
  const int n_seg = R * n_dst;
  GatherArgs a;
  a.out = agg; a.out_lo = agg_lo; a.ld_out = ld_agg; a.n_out_rows = n_dst;
  a.src = x; a.ld_src = D;
  // with no edges at all the (empty) support / end-point arrays may be null; they are never read, but the
  // kernel variant is chosen by pointer, so keep the weighted variant selected
  a.w = support ? support : reinterpret_cast<const float *>(cat_indptr);
  a.idx = end_points ? end_points : cat_indptr;
  a.indptr = cat_indptr; a.F = D; a.req = SG_REQ_WRITE;
  a.wsum = wsum; a.wsum_lo = wsum_lo; a.wsum_ld = wsum_ld;
  a.plan_chunk = plan_chunk;
  if (plan) {
    // partial scratch: rows of D floats followed by one weight-sum per row
    const size_t rows = sg_plan_partial_rows(n_seg, nnz, plan_chunk);
    a.partial = partial;
    a.partial_wsum = (wsum && partial) ? partial + rows * (size_t)D : nullptr;
  }
  return run_gather(a, 1, n_seg, nnz, plan, (cudaStream_t)stream);
}

int sg_multilink_agg_fwd(float *agg, float *wsum, const float *x, const float *support, const int32_t *end_points,
                         const int32_t *cat_indptr, int R, int n_dst, int n_nb, int nnz, int D, const void *plan,
                         int plan_chunk, float *partial, sg_stream_t stream) {
  return multilink_agg_fwd_impl(agg, nullptr, R * D, wsum, nullptr, R, x, support, end_points, cat_indptr, R, n_dst,
                                n_nb, nnz, D, plan, plan_chunk, partial, stream);
}

int sg_multilink_agg_fwd_split(float *agg_hi, float *agg_lo, int ld_agg, const float *x, const float *support,
                               const int32_t *end_points, const int32_t *cat_indptr, int R, int n_dst, int n_nb,
                               int nnz, int D, const void *plan, int plan_chunk, float *partial, sg_stream_t stream) {
  SG_REQUIRE(ld_agg >= R * D + R && (ld_agg & 3) == 0, "sg_multilink_agg_fwd_split: ld_agg must be a multiple of 4 and >= R*D + R");
  // agg_lo == NULL: the same [agg | wsum] row layout, plain fp32 (the in-kernel-split GEMM takes it as it is)
  return multilink_agg_fwd_impl(agg_hi, agg_lo, ld_agg, agg_hi + (size_t)R * D, agg_lo ? agg_lo + (size_t)R * D : nullptr, ld_agg, x, support,
                                end_points, cat_indptr, R, n_dst, n_nb, nnz, D, plan, plan_chunk, partial, stream);
}

int sg_multilink_agg_bwd(float *gx, const float *gagg, const float *t_w, const int32_t *t_src,
                         const int32_t *t_indptr, int R, int n_dst, int n_nb, int nnz, int D, int req,
                         const void *t_plan, int plan_chunk, float *partial, sg_stream_t stream) {
  SG_REQUIRE(valid_req(req), "sg_multilink_agg_bwd: bad req %d", req);
  if (req == SG_REQ_NULL) return SG_OK;
  SG_REQUIRE(R > 0 && n_dst >= 0 && n_nb >= 0 && nnz >= 0 && D > 0, "sg_multilink_agg_bwd: bad sizes");
  if (n_nb == 0) return SG_OK;
  SG_REQUIRE(gx && t_indptr && (nnz == 0 || (gagg && t_w && t_src)), "sg_multilink_agg_bwd: null pointer");
  GatherArgs a;
  a.out = gx; a.ld_out = D; a.n_out_rows = n_nb;
  a.src = gagg; a.ld_src = D;  // gagg [n_dst, R*D] viewed as [(n_dst*R), D]
  a.w = t_w; a.idx = t_src; a.indptr = t_indptr; a.F = D; a.req = req;
  a.plan_chunk = plan_chunk; a.partial = partial;
  return run_gather(a, 1, n_nb, nnz, t_plan, (cudaStream_t)stream);
}

int sg_multilink_agg_bwd_peer(float *const *stage_host, const int32_t *owner_lo_host, int n_targets, const float *gagg,
                              const float *t_w, const int32_t *t_src, const int32_t *t_indptr, int R, int n_dst,
                              int n_nb, int nnz, int D, const void *t_plan, int plan_chunk, float *partial,
                              sg_stream_t stream) {
  const int world = n_targets;
  SG_REQUIRE(world >= 1 && world <= SG_MAX_PEERS + 1, "sg_multilink_agg_bwd_peer: n_targets must be 1..%d", SG_MAX_PEERS + 1);
  SG_REQUIRE(R > 0 && n_dst >= 0 && n_nb >= 0 && nnz >= 0 && D > 0, "sg_multilink_agg_bwd_peer: bad sizes");
  SG_REQUIRE(stage_host && owner_lo_host && t_indptr && (nnz == 0 || (gagg && t_w && t_src)),
             "sg_multilink_agg_bwd_peer: null pointer");
  SG_REQUIRE(owner_lo_host[0] == 0 && owner_lo_host[world] == n_nb, "sg_multilink_agg_bwd_peer: ownership ranges must cover [0, n_nb)");
  if (n_nb == 0) return SG_OK;
  GatherArgs a;
  a.out = nullptr; a.ld_out = D; a.n_out_rows = n_nb;
  a.src = gagg; a.ld_src = D;
  a.w = t_w; a.idx = t_src; a.indptr = t_indptr; a.F = D; a.req = SG_REQ_WRITE;
  a.plan_chunk = plan_chunk; a.partial = partial;
  a.peer_world = world;
  for (int q = 0; q < world; ++q) {
    SG_REQUIRE(owner_lo_host[q] <= owner_lo_host[q + 1], "sg_multilink_agg_bwd_peer: ownership ranges must ascend");
    a.peer_lo[q] = owner_lo_host[q];
    a.peer_out[q] = stage_host[q];
  }
  a.peer_lo[world] = owner_lo_host[world];
  return run_gather(a, 1, n_nb, nnz, t_plan, (cudaStream_t)stream);
}

}  // extern "C"
