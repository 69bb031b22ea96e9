// TEST INFRASTRUCTURE ONLY — never linked or imported by the product path.
//
// Exposes the reference's host-side integer bookkeeping
// (/root/reference/GraphSampler/graph_sampler.{h,cpp}, compiled unmodified from where
// it lies with the two-header sparsehash shim in oracle/shim) behind a C ABI, so the
// device-side CSR bookkeeping (segment ids, per-rating-level split, support, fixed
// fan-out sampling, batch-edge removal) can be checked bit-exactly.
//
// Output buffers are caller-allocated; every function returns the number of elements
// written (or -1 if the caller's capacity is too small).
#include <cstring>
#include <vector>
#include "graph_sampler.h"

using namespace graph_sampler;

extern "C" {

// graph_sampler.cpp:378-391
int ref_gen_row_indices_by_indptr(const int* ind_ptr, int num, int nnz, int* out) {
  std::vector<int> rows;
  gen_row_indices_by_indptr(ind_ptr, num, nnz, &rows);
  std::memcpy(out, rows.data(), sizeof(int) * rows.size());
  return static_cast<int>(rows.size());
}

// graph_sampler.cpp:393-420
int ref_get_support(const int* row_degrees, const int* col_degrees, const int* ind_ptr,
                    const int* end_points, int num, int nnz, int symm, float* out) {
  std::vector<float> s;
  get_support(row_degrees, col_degrees, ind_ptr, end_points, num, nnz, symm != 0, &s);
  std::memcpy(out, s.data(), sizeof(float) * s.size());
  return static_cast<int>(s.size());
}

// graph_sampler.cpp:277-376. split_indices are positions into the unsplit nnz axis,
// written back-to-back (level 0 first) into out_indices; out_counts[v] is the length of
// level v's list; out_ind_ptrs is (val_num, node_num+1) row-major.
int ref_multi_link_split_by_value(const float* edge_values, const int* ind_ptr,
                                  const float* possible_values, int node_num, int nnz, int val_num,
                                  int* out_indices, int* out_counts, int* out_ind_ptrs) {
  std::vector<std::vector<int>> idx, ptr;
  multi_link_split_by_value(edge_values, ind_ptr, possible_values, node_num, nnz, val_num, &idx, &ptr);
  int w = 0;
  for (int v = 0; v < val_num; v++) {
    std::memcpy(out_indices + w, idx[v].data(), sizeof(int) * idx[v].size());
    out_counts[v] = static_cast<int>(idx[v].size());
    w += out_counts[v];
    if (static_cast<int>(ptr[v].size()) != node_num + 1) return -1;
    std::memcpy(out_ind_ptrs + static_cast<size_t>(v) * (node_num + 1), ptr[v].data(),
                sizeof(int) * (node_num + 1));
  }
  return w;
}

// graph_sampler.cpp:742-779 (GraphSampler::set_seed graph_sampler.h:176-202)
int ref_random_sample_fix_neighbor(int seed, const int* src_ind_ptr, const int* sel_indices,
                                   int sel_node_num, int neighbor_num, int cap,
                                   int* out_sampled, int* out_ind_ptr) {
  GraphSampler gs(seed);
  std::vector<int> sampled, ptr;
  gs.random_sample_fix_neighbor(src_ind_ptr, sel_indices, sel_node_num, neighbor_num, &sampled, &ptr);
  if (static_cast<int>(sampled.size()) > cap) return -1;
  std::memcpy(out_sampled, sampled.data(), sizeof(int) * sampled.size());
  std::memcpy(out_ind_ptr, ptr.data(), sizeof(int) * ptr.size());
  return static_cast<int>(sampled.size());
}

// graph_sampler.cpp:154-201 (the serial, order-defining variant)
int ref_remove_edges(const int* end_points, const float* values, const int* ind_ptr,
                     const int* row_indices, const int* col_indices, int row_num, int nnz,
                     int edge_num, int* out_end_points, float* out_values, int* out_ind_ptr) {
  std::vector<int> ep, ptr;
  std::vector<float> val;
  remove_edges(end_points, values, ind_ptr, row_indices, col_indices, row_num, nnz, edge_num,
               &ep, &val, &ptr);
  std::memcpy(out_end_points, ep.data(), sizeof(int) * ep.size());
  std::memcpy(out_values, val.data(), sizeof(float) * val.size());
  std::memcpy(out_ind_ptr, ptr.data(), sizeof(int) * ptr.size());
  return static_cast<int>(ep.size());
}

}  // extern "C"
