// TEST INFRASTRUCTURE ONLY — never linked or imported by the product path.
//
// Compiles the reference authors' own plain-pointer CPU loops straight out of
// /root/reference/seg_ops_cuda/seg_ops.cu (textually #included from where it lies;
// nothing is copied into this repo) and exports them behind a C ABI so that
// oracle/seg_ops_oracle.c (our restatement) can be pinned against the real thing.
//
//   SegTakeKCorrCPU<OP>             seg_ops.cu:788-806   (= weight-grad of seg_weighted_pool)
//   SegTakeKCorrCPUBackwardEmbed1   seg_ops.cu:808-824   (= seg_weighted_pool forward)
//   SegTakeKCorrCPUBackwardEmbed2   seg_ops.cu:826-844   (= seg_weighted_pool data-grad)
//   SegPoolCPU<type>                seg_ops.cu:1048-1096
//   SegPoolBackwardCPU<type>        seg_ops.cu:1104-1129
//
// The GPU half of that file is compiled too (nvcc has to parse it) but is dead code
// here: CUB is replaced by inert stubs from oracle/shim (see the comment there) and
// the prototype's main() is renamed away.
#include <cfloat>
#define main stargcn_ref_prototype_main
#include STARGCN_REF_SEG_OPS_CU
#undef main

extern "C" {

void ref_take_k_corr(float* dst, const float* embed1, const float* embed2, const int* ids,
                     const int* indptr, int K, int node_num, int nb_num, int nnz, int F) {
  SegTakeKCorrCPU<mul>(dst, embed1, embed2, ids, indptr, K, node_num, nb_num, nnz, F);
}

void ref_weighted_pool_fwd(float* dst, const float* w, const float* data, const int* ids,
                           const int* indptr, int K, int node_num, int nb_num, int nnz, int F) {
  SegTakeKCorrCPUBackwardEmbed1(dst, w, data, ids, indptr, K, node_num, nb_num, nnz, F);
}

void ref_weighted_pool_bwd_data(float* dst, const float* w, const float* gout, const int* ids,
                                const int* indptr, int K, int node_num, int nb_num, int nnz, int F) {
  SegTakeKCorrCPUBackwardEmbed2(dst, w, gout, ids, indptr, K, node_num, nb_num, nnz, F);
}

// pool_type: 0 sum, 1 mean, 2 max (SegReduceType, seg_ops.cu:34)
void ref_seg_pool_fwd(float* dst, int* dst_index, const float* data, const int* indices,
                      const int* indptr, int B, int seg_num, int F, int total, int nnz, int pool_type) {
  if (pool_type == 0) SegPoolCPU<SegReduceType::kSum>(dst, dst_index, data, indices, indptr, B, seg_num, F, total, nnz);
  else if (pool_type == 1) SegPoolCPU<SegReduceType::kMean>(dst, dst_index, data, indices, indptr, B, seg_num, F, total, nnz);
  else SegPoolCPU<SegReduceType::kMax>(dst, dst_index, data, indices, indptr, B, seg_num, F, total, nnz);
}

void ref_seg_pool_bwd(float* dst, const float* gout, const int* out_index, const int* indices,
                      const int* indptr, int B, int seg_num, int F, int total, int nnz, int pool_type) {
  if (pool_type == 0) SegPoolBackwardCPU<SegReduceType::kSum>(dst, gout, out_index, indices, indptr, B, seg_num, F, total, nnz);
  else if (pool_type == 1) SegPoolBackwardCPU<SegReduceType::kMean>(dst, gout, out_index, indices, indptr, B, seg_num, F, total, nnz);
  else SegPoolBackwardCPU<SegReduceType::kMax>(dst, gout, out_index, indices, indptr, B, seg_num, F, total, nnz);
}

}  // extern "C"
