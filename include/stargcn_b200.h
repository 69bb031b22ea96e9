/*
 * stargcn_b200 — C ABI of the Blackwell-native STAR-GCN aggregation hot path.
 *
 * Drop-in boundary for the reference's MXNet operator plug-in
 * (/root/reference/seg_ops_cuda/mxnet_op/seg_op.{h,cc,cu}).  The reference registers
 * FCompute<gpu> functions that unpack TBlobs into (pointer, shape) pairs and call
 * `seg_op::*Impl(dst, ..., req, ctx, stream)` (seg_op.h:31-177); each entry point below is
 * what such an FCompute body would call instead.  Plain pointers and sizes only — no MXNet,
 * torch or C++ types cross this boundary.
 *
 * Rules common to every entry point
 *   - all pointers are DEVICE pointers unless the name says `_host`; float tensors are dense
 *     row-major fp32, index tensors int32 (the only dtypes the reference accepts:
 *     seg_op.h:232-233,410-413,530-532)
 *   - the caller owns every buffer including scratch (`ws`), exactly as MXNet owns its
 *     kTempSpace (seg_op.cc:365-368); the library never allocates device memory, never
 *     synchronises `stream`, and keeps no state between calls except the thread-local
 *     error string
 *   - `req` is MXNet's OpReqType (SG_REQ_*): NULL returns immediately, WRITE overwrites the
 *     whole destination (empty segments become 0), ADD accumulates into it
 *     (seg_op.cc:188-196; exercised by test_seg_ops.py:130,152-154)
 *   - return value: SG_OK or an SG_ERR_* code; sg_last_error() describes the last failure
 *     on the calling thread.  Nothing ever calls exit() (the reference does: seg_op.cu:14-18)
 */
#ifndef STARGCN_B200_H_
#define STARGCN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *sg_stream_t; /* a cudaStream_t */

enum { SG_OK = 0, SG_ERR_INVALID = 1, SG_ERR_CUDA = 2, SG_ERR_WORKSPACE = 3 };
enum { SG_REQ_NULL = 0, SG_REQ_WRITE = 1, SG_REQ_ADD = 3 };      /* mxnet::OpReqType */
enum { SG_POOL_SUM = 0, SG_POOL_MEAN = 1, SG_POOL_MAX = 2 };     /* SegReduceType, seg_op.h:20 */
enum { SG_REDUCE_SUM = 0, SG_REDUCE_MAX = 2, SG_REDUCE_MIN = 3 };
enum { SG_BCAST_ADD = 0, SG_BCAST_MUL = 1, SG_BCAST_TO = 2, SG_BCAST_SUB = 3, SG_BCAST_DIV = 4 };
#define SG_MAX_PEERS 8 /* ranks of one NVSwitch box */
enum { SG_ACT_IDENTITY = 0, SG_ACT_LEAKY = 1, SG_ACT_RELU = 2 }; /* common.py:32-57 */

const char *sg_last_error(void);
int sg_abi_version(void);
/* Number of kernels this library has launched (process-wide) since the last reset. */
long long sg_launch_count(void);
void sg_launch_count_reset(void);
/* Development only (tools/, A/B measurements): set a process-wide tuning option; 0 is always the shipped
 * behaviour and the product path never calls this.  Options: 0 gather variant (1 = bulk-copy staged segments),
 * 1 gather blocks per SM, 2 GEMM TMEM hand-back arrival (1 = release), 3 GEMM chain length in k-blocks, 4 threads per block of the fast gather launches, 5 quarter-blocks per SM of the all-gather store kernel,
 * 8 GEMM wait-cycle trace (sg_gemm_trace_read).  Nothing in the library reads the environment. */
int sg_dev_option(int which, int value);

/* ------------------------------------------------------------------------------------------
 * A7  Segment bookkeeping (bit-exact integer work)
 * replaces GetSegId::compute (seg_op.cu:91-110), the CPU seg_ids loop (seg_op.cc:226-231)
 * and gen_row_indices_by_indptr (GraphSampler/graph_sampler.cpp:378-391)
 * ---------------------------------------------------------------------------------------- */
int sg_seg_ids(int32_t *seg_ids /*nnz*/, const int32_t *indptr /*n_seg+1*/, int n_seg, int nnz,
               sg_stream_t stream);

/* Stable transpose of a CSR pattern: for every destination row n (a value of `indices`) the
 * nnz positions p that point at it, in ascending p.  Built ONCE per sampled plan and reused by
 * every backward call, where the reference re-runs iota + radix sort + scan on every call
 * (compute_grad_embed2, seg_op.cu:882-926; SegPool compute_grad_data, seg_op.cu:1250-1283).
 *   t_indptr (n_nb+1), t_perm (nnz): original position p, t_seg (nnz): segment owning p */
size_t sg_csr_transpose_ws_bytes(int n_seg, int n_nb, int nnz);
int sg_csr_transpose(int32_t *t_indptr, int32_t *t_perm, int32_t *t_seg, const int32_t *indices,
                     const int32_t *indptr, int n_seg, int n_nb, int nnz, void *ws, size_t ws_bytes,
                     sg_stream_t stream);

/* Segment schedule: cuts every segment of a CSR into work items of at most `chunk` edges so
 * that heavy-tailed degree distributions load-balance across the 148 SMs.  Opaque device
 * buffer of sg_plan_bytes(); depends on indptr only.  sg_plan_partial_rows() is the number
 * of scratch rows (each F floats) the gather kernels need for split segments.  Entry points
 * that accept a plan also take the `plan_chunk` it was built with (0 when plan is NULL). */
size_t sg_plan_bytes(int n_seg, int nnz, int chunk);
size_t sg_plan_partial_rows(int n_seg, int nnz, int chunk);
int sg_plan_build(void *plan, size_t plan_bytes, const int32_t *indptr, int n_seg, int nnz,
                  int chunk, sg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A2  seg_weighted_pool forward
 * replaces SegWeightedPoolForward<gpu> -> SegTakeKCorrBackwardEmbed1Impl<gpu>
 * (seg_op.h:460-476, seg_op.cu:956-979, kernel seg_op.cu:682-722)
 *   dst[k,s,:] (=|+=) sum_{p in seg s} weights[k,p] * data[k, indices[p], :]
 * `plan` may be NULL (one work item per segment); `partial` needs
 * K * sg_plan_partial_rows() * F floats when a plan is given.
 * ---------------------------------------------------------------------------------------- */
int sg_weighted_pool_fwd(float *dst /*K,n_seg,F*/, const float *data /*K,n_nb,F*/,
                         const float *weights /*K,nnz*/, const int32_t *indices /*nnz*/,
                         const int32_t *indptr /*n_seg+1*/, int K, int n_seg, int n_nb, int nnz, int F,
                         int req, const void *plan, int plan_chunk, float *partial, sg_stream_t stream);

/* A3  data-gradient of seg_weighted_pool
 * replaces _backward_seg_take_k_corr_embed2 -> SegTakeKCorrBackwardEmbed2Impl<gpu>
 * (seg_op.cc:700-712, seg_op.cu:981-1006, compute_grad_embed2 seg_op.cu:882-926)
 *   gdata[k, indices[p], :] (=|+=) weights[k,p] * gout[k, seg(p), :]
 * computed as a gather over the transposed pattern (no atomics, no per-call sort).
 * `t_plan` is a plan built on t_indptr (or NULL). */
int sg_weighted_pool_bwd_data(float *gdata /*K,n_nb,F*/, const float *gout /*K,n_seg,F*/,
                              const float *weights /*K,nnz*/, const int32_t *t_indptr,
                              const int32_t *t_perm, const int32_t *t_seg, int K, int n_seg, int n_nb,
                              int nnz, int F, int req, const void *t_plan, int plan_chunk, float *partial,
                              sg_stream_t stream);

/* A4  seg_take_k_corr (inner product) = weight-gradient of seg_weighted_pool
 * replaces SegTakeKCorrImpl<gpu> (seg_op.cu:928-954, kernel seg_op.cu:573-664)
 *   dst[k,p] (=|+=) <embed1[k, seg(p), :], embed2[k, indices[p], :]> */
int sg_take_k_corr(float *dst /*K,nnz*/, const float *embed1 /*K,n_node,F*/,
                   const float *embed2 /*K,n_nb,F*/, const int32_t *indices, const int32_t *indptr,
                   int K, int n_node, int n_nb, int nnz, int F, int req, sg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A5  seg_pool (sum | mean | max+argmax) forward / backward
 * replaces SegPoolForward / SegSumMeanPoolBackward / SegMaxPoolBackward
 * (seg_op.h:542-619, seg_op.cu:1057-1135,1171-1215,1286-1350)
 * argmax holds the POSITION p on the nnz axis (seg_op.cc:282), -1 for an empty segment
 * whose value is 0; ties keep the first position.  Forward has no ADD mode (seg_op.cc:252).
 * ---------------------------------------------------------------------------------------- */
int sg_seg_pool_fwd(float *dst /*B,n_seg,F*/, int32_t *argmax /*B,n_seg,F or NULL*/,
                    const float *data /*B,n_nb,F*/, const int32_t *indices, const int32_t *indptr,
                    int B, int n_seg, int n_nb, int nnz, int F, int pool_type, const void *plan,
                    int plan_chunk, float *partial, sg_stream_t stream);
int sg_seg_pool_bwd(float *gdata /*B,n_nb,F*/, const float *gout /*B,n_seg,F*/,
                    const int32_t *argmax /*max only*/, const int32_t *indptr, const int32_t *t_indptr,
                    const int32_t *t_perm, const int32_t *t_seg, int B, int n_seg, int n_nb, int nnz,
                    int F, int pool_type, int req, const void *t_plan, int plan_chunk, float *partial,
                    sg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A6  contiguous segment ops over a (B, nnz) array
 * replaces SegReduceImpl<gpu> (seg_op.cu:174-269) and SegBroadcastBinaryImpl<gpu>
 * (seg_op.cu:282-330); seg_sum's gradient is SG_BCAST_TO (seg_op.cc:370-379)
 * ---------------------------------------------------------------------------------------- */
int sg_seg_reduce(float *dst /*B,n_seg*/, const float *data /*B,nnz*/, const int32_t *indptr, int B,
                  int nnz, int n_seg, int reduce_type, int req, sg_stream_t stream);
int sg_seg_broadcast_binary(float *dst /*B,nnz*/, const float *lhs /*B,nnz or NULL for TO*/,
                            const float *rhs /*B,n_seg*/, const int32_t *indptr, int B, int nnz,
                            int n_seg, int op, int req, sg_stream_t stream);
/* seg_softmax forward/backward (seg_op.cu:385-513) — one fused kernel each. */
int sg_seg_softmax_fwd(float *dst, const float *data, const int32_t *indptr, int B, int nnz, int n_seg,
                       sg_stream_t stream);
int sg_seg_softmax_bwd(float *dst, const float *ograd, const float *val, const int32_t *indptr, int B,
                       int nnz, int n_seg, int req, sg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A1  Fused multi-relation aggregation (all R rating levels in one launch, aggregate-first)
 * replaces the per-level FullyConnected + seg_weighted_pool + add_n/concat loop of
 * MultiLinkGCNAggregator.hybrid_forward (mxgraph/layers/aggregators.py:133-159), using
 *   sum_p s[p] (W_r x[e[p]] + b_r) = W_r (sum_p s[p] x[e[p]]) + b_r sum_p s[p].
 * The R CSRs are concatenated relation-major: segment id = r * n_dst + i.
 *   agg [n_dst, R*D]   agg[i, r*D:(r+1)*D] = sum_{p in seg(r,i)} support[p] * x[end_points[p], :]
 *   wsum[n_dst, R]     wsum[i, r]          = sum_{p in seg(r,i)} support[p]
 * Backward (data-gradient): the same gather over the transposed pattern, reading
 * gagg viewed as [(n_dst*R), D] rows:  gx[n,:] = sum_q t_w[q] * gagg_row[t_src[q]].
 * ---------------------------------------------------------------------------------------- */
int sg_multilink_agg_fwd(float *agg, float *wsum, const float *x /*n_nb,D*/,
                         const float *support /*nnz*/, const int32_t *end_points /*nnz*/,
                         const int32_t *cat_indptr /*R*n_dst+1*/, int R, int n_dst, int n_nb, int nnz,
                         int D, const void *plan, int plan_chunk, float *partial, sg_stream_t stream);
/* Same aggregation written as the pre-split A operand of sg_gemm_tf32x3: row i of agg_hi / agg_lo
 * (ld_agg floats, ld_agg >= R*D + R, multiple of 4) holds the R aggregated D-vectors followed by the
 * R support sums wsum[i, r] at column R*D + r, each value x stored as (tf32-exact hi, x - hi).
 * Columns beyond R*D + R are not written.  D must be 16, 32, 64 or 128.  agg_lo == NULL writes the same
 * row layout as plain fp32 into agg_hi (the raw A operand of sg_gemm_tf32x3). */
int sg_multilink_agg_fwd_split(float *agg_hi, float *agg_lo, int ld_agg, const float *x, const float *support,
                               const int32_t *end_points, const int32_t *cat_indptr, int R, int n_dst, int n_nb,
                               int nnz, int D, const void *plan, int plan_chunk, float *partial, sg_stream_t stream);
/* Per-plan preparation of the transposed operands from sg_csr_transpose() outputs:
 *   t_src[q] = i*R + r  for the segment s = r*n_dst + i owning position t_perm[q]
 *   t_w[q]   = support[t_perm[q]] */
int sg_multilink_transpose_finish(int32_t *t_src, float *t_w, const int32_t *t_perm,
                                  const int32_t *t_seg, const float *support, int R, int n_dst, int nnz,
                                  sg_stream_t stream);
/* Upload n arrays from PINNED host memory (device-accessible under unified addressing) into device buffers with
 * one kernel launch instead of n DMA copies — the 3*R per-level lists of a small plan (the reference uploads each
 * with its own nd.array call, layers.py:366-377).  dst_device / src_pinned_host / bytes are HOST arrays of n entries;
 * every segment 4-byte aligned and a multiple of 4 bytes; zero-length segments are skipped. */
int sg_upload_segments(void *const *dst_device, const void *const *src_pinned_host, const size_t *bytes, int n,
                       sg_stream_t stream);
int sg_multilink_agg_bwd(float *gx /*n_nb,D*/, const float *gagg /*n_dst,R*D*/, const float *t_w,
                         const int32_t *t_src, const int32_t *t_indptr, int R, int n_dst, int n_nb,
                         int nnz, int D, int req, const void *t_plan, int plan_chunk, float *partial,
                         sg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (e)  Node-partitioned aggregation over NVLink / NVSwitch peer memory (no reference counterpart: the reference is
 * single-device, experiments/STAR-GCN.py:32).  Every rank has mapped every rank's exchange buffer (symmetric memory);
 * the `*_host` arguments are HOST arrays of `world` (<= SG_MAX_PEERS) DEVICE pointers, entry q addressing rank q's memory.
 *   sg_peer_push_rows   all-gather / all-reduce input: store src[n_floats] to dst[q] for every q (one read, world stores)
 *   sg_peer_barrier     flags[q] = rank q's flag array (>= world words, zero-initialised); state = 2 device words
 *                       {epoch, error}: release-store the next epoch into flags[q][rank] for all q, acquire-spin until
 *                       flags[rank][q] reached it.  A wait longer than timeout_s sets state[1] = 1 + missing rank and
 *                       returns.  CUDA-graph replayable (the epoch is device state)
 *   sg_peer_reduce      out (=|+=) stage[0] + stage[1] + ... + stage[world-1] in rank order, slot q at
 *                       stage + q * slot_stride_floats — the local half of the reduce-scatter / all-reduce
 *   sg_multilink_agg_bwd_peer   sg_multilink_agg_bwd whose output row j is stored at stage[q] + (j - owner_lo[q]) * D
 *                       for the target q with owner_lo[q] <= j < owner_lo[q+1] (n_targets <= SG_MAX_PEERS + 1 ranges
 *                       covering [0, n_nb); empty ranges allowed): the exchange's transfer happens inside the gather.
 *                       Dense halo: rows = global neighbour ids, one target per rank (the reduce-scatter).  Sparse halo:
 *                       target 0 = the rank's own rows, then one range of halo slots per peer (the reverse all-to-all);
 *                       with R = 1, unit weights and one edge per row it is also the forward pack-and-push of the
 *                       deduplicated halo rows.  D in {16, 32, 64, 128}.
 * ---------------------------------------------------------------------------------------- */
int sg_peer_push_rows(float *const *dst_host, const float *src, long long n_floats, int world, sg_stream_t stream);
int sg_peer_barrier(uint32_t *const *flags_host, uint32_t *state, int rank, int world, double timeout_s,
                    sg_stream_t stream);
int sg_peer_reduce(float *out, const float *stage, long long n_floats, long long slot_stride_floats, int world,
                   int req, sg_stream_t stream);
int sg_multilink_agg_bwd_peer(float *const *stage_host, const int32_t *owner_lo_host, int n_targets, const float *gagg,
                              const float *t_w, const int32_t *t_src, const int32_t *t_indptr, int R, int n_dst,
                              int n_nb, int nnz, int D, const void *t_plan, int plan_chunk, float *partial,
                              sg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A1 (transform part)  fp32-accurate GEMM on tcgen05 tensor cores (3xTF32 split)
 * replaces the R FullyConnected calls + add_n of aggregators.py:141-159 (cuBLAS sgemm inside
 * MXNet) and their backward.  An operand is either pre-split (x = hi + lo, sg_split_tf32: pass both
 * pointers) or plain fp32 (pass it as *_hi with *_lo == NULL: the kernel splits every tile in shared
 * memory — half the operand bytes and no producer pass; B may be raw only when A is):
 *   mn_major == 0:  D[M,N] = A[M,K] . B[N,K]^T      A, B row-major, K contiguous (lda, ldb)
 *   mn_major == 1:  D[M,N] = A[K,M]^T . B[K,N]      A, B row-major, M / N contiguous (reduction over rows)
 * lda / ldb must be multiples of 4 floats and the operands 16-byte aligned (TMA).
 * epilogue 0: store; 1: leaky-ReLU with `slope` (0 = ReLU, 1 = identity).
 * splits > 1 partitions K over CTAs; partials go to split_ws (sg_gemm_split_ws_bytes) and are
 * summed in a fixed order, so results are bit-identical run to run.
 * ---------------------------------------------------------------------------------------- */
size_t sg_gemm_split_ws_bytes(int M, int N, int splits);
int sg_gemm_tf32x3(float *D, int ldd, const float *A_hi, const float *A_lo, int lda, const float *B_hi,
                   const float *B_lo, int ldb, int M, int N, int K, int mn_major, int epilogue, float slope,
                   const float *bias /*N or NULL, added before the activation*/, int splits, float *split_ws,
                   sg_stream_t stream);
/* hi = src with the 13 low mantissa bits cleared, lo = src - hi; optional transpose; the
 * destination has ld_dst >= columns and its padding is zero-filled. */
int sg_split_tf32(float *hi, float *lo, int ld_dst, const float *src, int rows, int cols, int ld_src,
                  int transpose, sg_stream_t stream);
/* gZ = gout * act'(Z) evaluated from the saved layer output (leaky slope; 0 = ReLU), written
 * pre-split and padded to ldz columns: the A operand of both backward GEMMs
 * (replaces MXNet's LeakyReLU backward, mxgraph/layers/common.py:46-47). */
int sg_act_bwd_split(float *gz_hi, float *gz_lo, int ldz, const float *gout, const float *out, int M, int U,
                     float slope, sg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * D1-D4  masked-embedding reconstruction decoder and rating head (experiments/STAR-GCN.py)
 * ---------------------------------------------------------------------------------------- */
/* D1  Net.get_embed (STAR-GCN.py:264-300): id' = noise ? noise[ids[i]] : ids[i];
 * out[i,:] = id' == -1 ? 0 : table[id',:];  eff_ids[i] = id' (kept for the backward scatter).
 * Replaces take + not_equal + mul + Embedding + mul (5 MXNet launches). */
int sg_masked_embed_fwd(float *out /*n,D*/, int32_t *eff_ids /*n*/, const float *table /*n_table,D*/,
                        const int32_t *ids /*n*/, const int32_t *noise /*n_table or NULL*/, int n, int n_table,
                        int D, sg_stream_t stream);
/* D3  loss = scale * sum_i sum_d (a[i,d] - b[i,d])^2 — mx.nd.mean(mx.nd.sum(mx.nd.square(gt - pred), -1))
 * with scale = 1/n (STAR-GCN.py:625) and gluon L2Loss(...).mean() with D = 1, scale = 0.5/n (:611-616).
 * Two-stage fixed-order reduction (bit-identical reruns); ws needs sg_reduce_ws_bytes() bytes. */
size_t sg_reduce_ws_bytes(void);
int sg_sq_err_fwd(float *loss /*1*/, const float *a, const float *b, long long n_elem, float scale, void *ws,
                  sg_stream_t stream);
/* ga = 2 * scale * gloss[0] * (a - b),  gb = -ga  (either may be NULL) */
int sg_sq_err_bwd(float *ga, float *gb, const float *a, const float *b, const float *gloss /*1, device*/,
                  long long n_elem, float scale, sg_stream_t stream);
/* D4  InnerProductLayer without mid map (mxgraph/layers/layers.py:217-222): out[i] = <a[i,:], b[i,:]> */
int sg_rowdot_fwd(float *out /*n*/, const float *a, const float *b, int n, int D, sg_stream_t stream);
int sg_rowdot_bwd(float *ga, float *gb, const float *gout /*n*/, const float *a, const float *b, int n, int D,
                  sg_stream_t stream);
/* Bias gradient of a Dense layer: out[c] = sum_r (x_hi[r,c] + x_lo[r,c]) over a pre-split operand
 * (x_lo may be NULL); fixed-order two-stage reduction; ws needs sg_colsum_ws_bytes(N) bytes. */
size_t sg_colsum_ws_bytes(int N);
int sg_colsum(float *out /*N*/, const float *x_hi, const float *x_lo, int M, int N, int ld, void *ws,
              sg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Next rows  device-side neighbour sampling, per-level split and support
 * replaces CSRMat.sample_neighbors (mxgraph/graph.py:677-748) and the GraphSampler functions it
 * calls; the whole-graph CSR (ind_ptr, end_points, values) stays resident on the device.
 * ---------------------------------------------------------------------------------------- */
/* get_support (GraphSampler/graph_sampler.cpp:393-420): sqrt(1/d_row/d_col) (symm) or 1/d_row;
 * 0 where a degree is 0.  Bit-exact with the host code (IEEE division / square root). */
int sg_csr_support(float *support /*nnz*/, const int32_t *row_degrees, const int32_t *col_degrees,
                   const int32_t *indptr, const int32_t *end_points, int n_rows, int nnz, int symm,
                   sg_stream_t stream);
/* scratch for the two calls below (n_sel rows, R levels; R = 1 for the count call) */
size_t sg_sampler_ws_bytes(int n_sel, int R);
/* GraphSampler::random_sample_fix_neighbor (graph_sampler.cpp:742-779), step 1:
 * dst_indptr[i+1] - dst_indptr[i] = neighbor_num < 0 ? deg : min(neighbor_num, deg) for row sel[i]
 * (sel == NULL: every row); dst_indptr has n_sel + 1 entries, the last one is the sampled nnz. */
int sg_sample_neighbors_count(int32_t *dst_indptr, const int32_t *src_indptr, const int32_t *sel, int n_sel,
                              int neighbor_num /* < 0: all, else <= 256 */, void *ws, sg_stream_t stream);
/* step 2: positions on the nnz axis of the source CSR.  Rows whose quota equals their degree get
 * p_begin..p_end-1 in order (bit-exact); the others a partial Fisher-Yates draw without replacement
 * (uniform_choice_range, graph_sampler.cpp:698-732) from a generator keyed by (seed, source row, draw). */
int sg_sample_neighbors_fill(int32_t *sampled /*nnz_s*/, const int32_t *dst_indptr, const int32_t *src_indptr,
                             const int32_t *sel, int n_sel, unsigned long long seed, sg_stream_t stream);
/* multi_link_split_by_value (graph_sampler.cpp:277-312) + the np.take calls of graph.py:725-745, written
 * relation-major (segment r * n_sel + i):
 *   cat_indptr [R*n_sel + 1]   ind_ptr of level r, row i = cat_indptr[r*n_sel + i] - cat_indptr[r*n_sel]
 *   split_index[nnz_s]         position on the SAMPLED axis of every output edge (may be NULL)
 *   ep_cat / sup_cat / val_cat end point index, support, edge value of every output edge (sup/val may be NULL)
 *   bad_flag                   set to 1 if an edge value is not in possible_values (the reference ASSERTs) */
int sg_multilink_split(int32_t *cat_indptr, int32_t *split_index, int32_t *ep_cat, float *sup_cat, float *val_cat,
                       int32_t *bad_flag, const float *values, const int32_t *end_points, const float *support,
                       const int32_t *sampled, const int32_t *dst_indptr, const float *possible_values, int R,
                       int n_sel, void *ws, sg_stream_t stream);

/* Next rows  batch-edge removal: remove_edges (GraphSampler/graph_sampler.cpp:154-201,
 * called from HeterGraph.remove_edges_by_id, mxgraph/graph.py:952-974, every training iteration at
 * experiments/STAR-GCN.py:595-600).  Two steps so the caller can size the outputs:
 *   count: marks every copy of each listed (row, col) pair in ws, writes the new ind_ptr (n_rows + 1;
 *          the last entry is the new nnz)
 *   fill:  stable compaction of end_points / values (values may be NULL) using the marks left in ws */
size_t sg_remove_edges_ws_bytes(int n_rows, int nnz);
int sg_remove_edges_count(int32_t *dst_indptr, const int32_t *indptr, const int32_t *end_points,
                          const int32_t *rm_rows, const int32_t *rm_cols, int n_rows, int nnz, int n_rm, void *ws,
                          sg_stream_t stream);
int sg_remove_edges_fill(int32_t *dst_end_points, float *dst_values, const int32_t *dst_indptr,
                         const int32_t *indptr, const int32_t *end_points, const float *values, int n_rows,
                         int nnz, const void *ws, sg_stream_t stream);
/* Edge weights of a STATIC relation-major plan over the whole graph when a batch of edges is masked out instead of
 * removed (replaces the per-iteration CSR rebuild of remove_edges + get_support, graph.py:952-974 /
 * graph_sampler.cpp:393-420): plan position q (base position split_index[q], row plan_row[q], column plan_col[q])
 * gets 0 if keep[base position] == 0, else 1/sqrt(d_row d_col) with d = degrees after the removal
 * (new_*_ptr: the prefix sums sg_remove_edges_count writes for this matrix / for the reverse matrix). */
int sg_masked_support(float *support /*nnz*/, const int32_t *keep, const int32_t *new_row_ptr, const int32_t *new_col_ptr,
                      const int32_t *split_index, const int32_t *plan_row, const int32_t *plan_col, int nnz, int symm,
                      sg_stream_t stream);
/* counts[b] = |{i : idx[i] == b}| (column degrees after a removal; np.bincount) */
int sg_bincount(int32_t *counts /*n_bins*/, const int32_t *idx, int n, int n_bins, sg_stream_t stream);

/* Next rows  unique + inverse in first-occurrence order: the serial unique_inverse of
 * GraphSampler/graph_sampler.h:510-534 behind merge_nodes (mxgraph/graph.py:142-163), which gen_plan uses to
 * turn sampled end-point ids into local row indices (mxgraph/layers/layers.py:308-334).
 *   uniq[0..n_unique)  distinct values in order of first appearance     (capacity n)
 *   inverse[i]         index into uniq of data[i]                        (n)
 *   n_unique           device scalar */
size_t sg_unique_inverse_ws_bytes(int n);
int sg_unique_inverse(int32_t *uniq, int32_t *inverse, int32_t *n_unique, const int32_t *data, int n, void *ws,
                      size_t ws_bytes, sg_stream_t stream);

/* Next rows  multi-tensor global-norm clip + Adam step, one launch over all parameters.
 * replaces params_clip_global_norm (mxgraph/utils.py:104-107 -> gluon.utils.clip_global_norm) and
 * gluon.Trainer('adam').step (experiments/STAR-GCN.py:552-553,630-632 -> mx adam_update).
 * Tensors are addressed through DEVICE arrays of device pointers; `work` is a device array of
 * (tensor index, chunk index) int32 pairs, one per sg_optim_chunk() elements of every tensor.
 *   out2[0] = sqrt(sum ||g||^2),  out2[1] = min(1, max_norm / (norm + 1e-8));  ws: n_work floats */
int sg_optim_chunk(void);
int sg_global_norm(float *out2, const float *const *grads, const long long *numels, const void *work, int n_work,
                   float max_norm, float *ws, sg_stream_t stream);
/* g = g * clip_out2[1] (written back when write_back_grad) ; g = g * rescale + wd * w ;
 * m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; w -= lr_t * m / (sqrt(v) + eps)
 * lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) is computed by the caller; clip_out2 may be NULL. */
int sg_multi_adam(float *const *params, float *const *grads, float *const *ms, float *const *vs,
                  const long long *numels, const void *work, int n_work, float lr_t, float beta1, float beta2,
                  float eps, float wd, float rescale, const float *clip_out2, int write_back_grad,
                  sg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Measurement aid (bench.py roofline; no reference counterpart): random-row gather ceiling.
 * `blocks` x 16 lane groups each read `reads_per_group` (multiple of 8) rows of 64 floats at hashed
 * positions of table[n_rows, 64] — the access shape of the aggregation gather with no index load,
 * weight or epilogue — and write one row: out[blocks * 16, 64].  Bytes moved =
 * blocks * 16 * reads_per_group * 256.
 * ---------------------------------------------------------------------------------------- */
/* Development: with dev option 8 (gemm_trace) set, the GEMM kernels accumulate the cycles the roles of one CTA pair
 * spend waiting; this copies the 8 counters to the host and clears them (synchronises the device). */
int sg_gemm_trace_read(unsigned long long *host8);
/* Measurement aid: TMA delivery rate to one SM — `blocks` CTAs each stream `iters` ring slots of `boxes`
 * [32 floats x 128 rows] boxes (16 KB each) of src[rows, K] (ld floats per row) through `stages` slots.
 * Bytes moved = blocks * |iters| * boxes * 16384; iters < 0: every CTA reads the SAME boxes (hot L2 lines, the
 * access pattern of a weight tile shared by all CTAs); stages + 100 * (P - 1): P producer warps share a stage's boxes. */
int sg_tma_probe(const float *src, int rows, int K, int ld, int stages, int boxes, int iters, int blocks,
                 sg_stream_t stream);
int sg_row_gather_probe(float *out, const float *table, int n_rows, int reads_per_group, int blocks,
                        unsigned seed, sg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STARGCN_B200_H_ */
