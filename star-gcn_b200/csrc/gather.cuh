// Argument block of the CSR gather-accumulate kernels (gather.cu).
#pragma once
#include "common.cuh"

namespace sg {

struct GatherArgs {
  // destination: row for segment `seg` starts at out_offset(seg) (see gather.cu):
  //   n_out_rows == n_seg : seg * ld_out
  //   otherwise (relation-major concatenated CSRs, seg = r*n_out_rows + i): i*ld_out + r*F
  float *out = nullptr;
  long long out_batch_stride = 0;
  int ld_out = 0;
  int n_out_rows = 0;
  // seg / n_out_rows without an integer division (filled by run_gather; see seg_rel in gather.cu)
  uint32_t div_magic = 0;
  int div_shift = 0;
  // gathered matrix
  const float *src = nullptr;
  long long src_batch_stride = 0;
  int ld_src = 0;
  // per-edge weight (nullptr: 1.0) read at position p, or at perm[p] when perm != nullptr
  const float *w = nullptr;
  long long w_batch_stride = 0;
  const int32_t *perm = nullptr;
  // when set, the weight of edge p is 1 / (inv_len_indptr[id+1] - inv_len_indptr[id]), id = idx[p]
  // (seg_pool 'avg' backward: every contribution is divided by the length of ITS segment)
  const int32_t *inv_len_indptr = nullptr;
  // CSR pattern
  const int32_t *idx = nullptr;
  const int32_t *indptr = nullptr;
  int n_seg = 0;
  int F = 0;
  // schedule (filled by run_gather from the opaque plan buffer)
  const PlanHeader *hdr = nullptr;
  const int4 *items = nullptr;
  const int4 *longs = nullptr;
  int plan_chunk = 0;
  float *partial = nullptr;
  long long partial_batch_stride = 0;
  // optional per-segment sum of weights: wsum[i * wsum_ld + r] for segment r * n_out_rows + i
  float *wsum = nullptr;
  int wsum_ld = 0;
  float *partial_wsum = nullptr;
  // when set, results are stored pre-split for the 3xTF32 GEMM: out/wsum receive the TF32-exact
  // high part, out_lo/wsum_lo the remainder (same offsets)
  float *out_lo = nullptr;
  float *wsum_lo = nullptr;
  int req = SG_REQ_WRITE;
  int mean = 0;  // divide by the segment length (seg_pool 'avg')
  // peer_world > 0 (node-partitioned exchange, peer.cu): output row j belongs to TARGET q with peer_lo[q] <= j <
  // peer_lo[q+1] and is stored at peer_out[q] + (j - peer_lo[q]) * ld_out — a buffer of another rank reached over
  // NVLink (or a local one); `out` is unused.  Up to SG_MAX_PEERS + 1 targets (the rank's own rows + one range per
  // peer in the sparse-halo layout).  Write semantics, n_out_rows == n_seg.
  int peer_world = 0;
  int peer_lo[SG_MAX_PEERS + 2] = {0};
  float *peer_out[SG_MAX_PEERS + 1] = {nullptr};
};

int run_gather(GatherArgs a, int K, int n_seg, int nnz, const void *plan, cudaStream_t st);

}  // namespace sg
