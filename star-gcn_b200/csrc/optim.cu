// Multi-tensor global-norm clip + Adam step: one launch over ALL parameters.
//   params_clip_global_norm  mxgraph/utils.py:104-107 -> gluon.utils.clip_global_norm:
//        norm = sqrt(sum_t ||g_t||^2);  scale = max_norm / (norm + 1e-8);  g_t *= scale if scale < 1
//   gluon.Trainer('adam').step(1.0)  experiments/STAR-GCN.py:552-553,632 -> mx.optimizer.Adam / adam_update:
//        lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)         (host side)
//        g = g * rescale + wd * w;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  w -= lr_t * m / (sqrt(v) + eps)
// The reference launches ~6 MXNet ops per parameter per iteration (40+ parameters); here the tensors
// are addressed through device-resident pointer tables and a (tensor, chunk) work list built once.
#include "common.cuh"

namespace sg {

constexpr int kOptChunk = 4096;  // elements per work item
constexpr int kOptThreads = 256;

__device__ __forceinline__ float block_sum_opt(float v) {
  __shared__ float part[kOptThreads / 32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) part[wid] = v;
  __syncthreads();
  float t = 0.f;
  if (wid == 0) {
    t = lane < kOptThreads / 32 ? part[lane] : 0.f;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  }
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kOptThreads) multi_sqnorm_kernel(float *__restrict__ partial, const float *const *__restrict__ grads,
                                                                   const long long *__restrict__ numels,
                                                                   const int2 *__restrict__ work) {
  const int2 wk = work[blockIdx.x];
  const float *g = grads[wk.x];
  const long long lo = (long long)wk.y * kOptChunk, hi = min(lo + kOptChunk, numels[wk.x]);
  float acc = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += kOptThreads) {
    const float x = g[i];
    acc = fmaf(x, x, acc);
  }
  const float s = block_sum_opt(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[0] = norm, out[1] = scale applied to every gradient (1 when the norm is within max_norm)
__global__ void __launch_bounds__(kOptThreads) norm_finish_kernel(float *__restrict__ out, const float *__restrict__ partial,
                                                                  int n_partial, float max_norm) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_partial; i += kOptThreads) acc += partial[i];
  const float s = block_sum_opt(acc);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(s);
    const float scale = max_norm / (norm + 1e-8f);
    out[0] = norm;
    out[1] = scale < 1.f ? scale : 1.f;
  }
}

__global__ void __launch_bounds__(kOptThreads) multi_adam_kernel(float *const *__restrict__ params, float *const *__restrict__ grads,
                                                                 float *const *__restrict__ ms, float *const *__restrict__ vs,
                                                                 const long long *__restrict__ numels,
                                                                 const int2 *__restrict__ work, float lr_t, float beta1,
                                                                 float beta2, float eps, float wd, float rescale,
                                                                 const float *__restrict__ clip_scale, int write_back_grad) {
  const int2 wk = work[blockIdx.x];
  float *w = params[wk.x], *g = grads[wk.x], *m = ms[wk.x], *v = vs[wk.x];
  const long long lo = (long long)wk.y * kOptChunk, hi = min(lo + kOptChunk, numels[wk.x]);
  const float cs = clip_scale ? clip_scale[1] : 1.f;
  for (long long i = lo + threadIdx.x; i < hi; i += kOptThreads) {
    float gi = g[i] * cs;
    if (write_back_grad && clip_scale) g[i] = gi;  // clip_global_norm rescales the gradient arrays in place
    const float wi = w[i];
    gi = gi * rescale + wd * wi;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    w[i] = wi - lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_optim_chunk(void) { return kOptChunk; }

int sg_global_norm(float *out2, const float *const *grads, const long long *numels, const void *work, int n_work,
                   float max_norm, float *ws, sg_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_REQUIRE(n_work >= 0, "sg_global_norm: negative size");
  SG_REQUIRE(out2 && ws && (n_work == 0 || (grads && numels && work)), "sg_global_norm: null pointer");
  if (n_work > 0) {
    multi_sqnorm_kernel<<<n_work, kOptThreads, 0, st>>>(ws, grads, numels, static_cast<const int2 *>(work));
    SG_LAUNCHED("multi_sqnorm_kernel");
  }
  norm_finish_kernel<<<1, kOptThreads, 0, st>>>(out2, ws, n_work, max_norm);
  SG_LAUNCHED("norm_finish_kernel");
  return SG_OK;
}

int sg_multi_adam(float *const *params, float *const *grads, float *const *ms, float *const *vs, const long long *numels,
                  const void *work, int n_work, float lr_t, float beta1, float beta2, float eps, float wd, float rescale,
                  const float *clip_out2, int write_back_grad, sg_stream_t stream) {
  SG_REQUIRE(n_work >= 0, "sg_multi_adam: negative size");
  if (n_work == 0) return SG_OK;
  SG_REQUIRE(params && grads && ms && vs && numels && work, "sg_multi_adam: null pointer");
  multi_adam_kernel<<<n_work, kOptThreads, 0, (cudaStream_t)stream>>>(params, grads, ms, vs, numels, static_cast<const int2 *>(work),
                                                                     lr_t, beta1, beta2, eps, wd, rescale, clip_out2, write_back_grad);
  SG_LAUNCHED("multi_adam_kernel");
  return SG_OK;
}

}  // extern "C"
