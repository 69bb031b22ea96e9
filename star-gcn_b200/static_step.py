"""A whole STAR-GCN training iteration as ONE CUDA graph ("the last host round trip removed").

The reference builds, every iteration, a new graph without the batch's rating edges (single-threaded CSR rebuild of
both directions, experiments/STAR-GCN.py:595-600 -> mxgraph/graph.py:952-974), samples the neighbourhoods of the
batch nodes, merges ids into local index spaces (gen_plan, layers.py:260-337) and re-uploads every index array
(layers.py:366-377).  All shipped configurations use the FULL neighbourhood (NUM_NEIGHBORS = -1), and two stacked
blocks over a rating graph reach practically every node — so the plan of an iteration is, up to which rows are
read at the end, the whole graph minus the batch edges.  ``StaticGraphStep`` therefore keeps ONE relation-major
plan of the whole training graph per direction resident and changes only what changes:

  * batch edges are MASKED, not removed: their weight becomes 0 (``fma(0, x, acc) == acc``: bit-neutral) and every
    other weight is re-evaluated as 1/sqrt(d_row d_col) with the degrees of the graph without the batch edges —
    exactly what ``remove_edges`` + ``get_support`` (graph_sampler.cpp:154-201, 393-420) would give
    (``sg_remove_edges_count`` marks and counts, ``sg_masked_support`` writes the plan's weights);
  * every block runs over ALL nodes of both types (local index == row index: no id merging, no ``take``), the
    rating head and the reconstruction decoder read the rows of the batch / recon nodes at the end.

Every shape is then fixed by (graph, batch size, number of recon nodes): mask update, embedding lookup, both blocks
forward and backward, decoder and losses are captured once and replayed; per iteration the host copies the batch
(pairs, ratings, noise tables, recon ids) into static buffers and launches one graph (+ the two launches of the
fused clip + Adam).  Outputs equal ``StarGCN.forward`` on the edge-removed graph to fp32 rounding (work items cut
long segments at different places once zero-weight edges sit between the kept ones; tests/test_static_step_gpu.py).
"""
import torch

from . import _lib, decoder, seg_op
from ._lib import check
from .devgraph import DeviceHeterGraph, _i32
from .graph import MultiLinkCSR
from .seg_op import _bytes, _p, _stream


class _Direction:
    """Static whole-graph plan of one (src, dst) matrix + what the per-iteration weight update needs."""

    def __init__(self, mat):
        g = mat.csr
        self.mat, self.g = mat, g
        dev = g.device
        sampled, dst_indptr, n_sel = g.sample_positions(None, -1, seed=0)          # every row, every edge, in order
        cat_indptr, ep_cat, sup_cat, split_index, _ = g.split(sampled, dst_indptr, n_sel, want_index=True)
        self.base_pos = sampled[split_index.long()].contiguous()                    # plan position -> base CSR position
        seg = seg_op.seg_ids(cat_indptr, g.nnz)
        self.plan_row = (seg % g.n_rows).to(torch.int32).contiguous()
        self.csr = MultiLinkCSR.from_device(ep_cat, sup_cat.clone(), cat_indptr, g.R, g.n_rows, g.n_cols)
        lib = _lib.load()
        self.ws = _bytes(lib.sg_remove_edges_ws_bytes(g.n_rows, g.nnz), dev)        # keep flags (int32 per edge) first
        self.new_ptr = torch.empty(g.n_rows + 1, dtype=torch.int32, device=dev)

    def mark(self, rm_rows, rm_cols):
        g = self.g
        check(_lib.load().sg_remove_edges_count(_p(self.new_ptr), _p(g.ind_ptr), _p(g.end_points), _p(rm_rows), _p(rm_cols),
                                                g.n_rows, g.nnz, rm_rows.numel(), _p(self.ws), _stream()),
              "sg_remove_edges_count")

    def reweigh(self, reverse, symm):
        g = self.g
        check(_lib.load().sg_masked_support(_p(self.csr.support), _p(self.ws), _p(self.new_ptr), _p(reverse.new_ptr),
                                            _p(self.base_pos), _p(self.plan_row), _p(self.csr.end_points), g.nnz, int(bool(symm)),
                                            _stream()), "sg_masked_support")


class StaticGraphStep:
    """``step = StaticGraphStep(model, graph, batch_size, n_recon); loss = step(pairs, ratings, noise, recon_ids)``.

    model       :class:`stargcn_b200.model.StarGCN` (bipartite user / item, one HeterGCNLayer per block, materialised)
    graph       :class:`DeviceHeterGraph` of the WHOLE training graph (batch edges still inside)
    batch_size  rating pairs per iteration (fixed: one captured graph per shape)
    n_recon     {node type: number of nodes to reconstruct per iteration}
    ``__call__`` takes host (numpy / pinned) or device arrays: pairs (2, B) node ids, ratings (B,), the noise tables
    {type: (N_type,) int32, -1 = masked to the zero vector} and recon ids {type: (n_recon[type],)}; it returns the
    loss as a device scalar and leaves the parameter gradients in ``.grad`` (call the optimiser afterwards)."""

    def __init__(self, model, graph, batch_size, n_recon, symm=True, rating_mean=0.0, rating_std=1.0, recon_lambda=0.1):
        if not isinstance(graph, DeviceHeterGraph):
            raise TypeError("StaticGraphStep needs a DeviceHeterGraph")
        self.model, self.graph, self.symm = model, graph, symm
        self.mean, self.std, self.lam = float(rating_mean), float(rating_std), float(recon_lambda)
        self.user, self.item = model._name_user, model._name_item
        dev = graph.device
        self.dev = dev
        u, i = self.user, self.item
        self.dirs = {(u, i): _Direction(graph[u, i]), (i, u): _Direction(graph[i, u])}
        for d in self.dirs.values():
            d.csr.keep_transpose_scratch = True       # the weights change every iteration, the pattern never does
            d.csr.prepare(backward=True)
        if not (torch.equal(graph[u, i].col_ids, graph[i, u].row_ids) and torch.equal(graph[i, u].col_ids, graph[u, i].row_ids)):
            raise ValueError("the columns of each direction must list the nodes in the order of the other direction's rows")
        self.all_ids = {u: graph[u, i].row_ids, i: graph[i, u].row_ids}
        B = int(batch_size)
        self.B = B
        self.pairs = torch.zeros((2, B), dtype=torch.int32, device=dev)
        self.ratings = torch.zeros(B, dtype=torch.float32, device=dev)
        self.noise = {k: torch.arange(model.embed_layers[k].weight.shape[0], dtype=torch.int32, device=dev) for k in (u, i)}
        self.recon = {k: torch.zeros(int(n), dtype=torch.int32, device=dev) for k, n in n_recon.items()}
        self.loss = None
        self._graph = None
        # the two node types are independent inside a block: their layers run as two branches (two streams eagerly,
        # parallel branches of the captured graph) — these graphs are chains of tiny kernels, latency-bound
        self._streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

    # ---- the captured body ----
    def _body(self):
        m, u, i = self.model, self.user, self.item
        d_ui, d_iu = self.dirs[(u, i)], self.dirs[(i, u)]
        rows_u = self.graph[u, i].rows_of(self.pairs[0]).contiguous()
        rows_i = self.graph[i, u].rows_of(self.pairs[1]).contiguous()
        # batch edges out of both directions: mark + degrees, then the plans' weights and transposed weights
        d_ui.mark(rows_u, rows_i)
        d_iu.mark(rows_i, rows_u)
        d_ui.reweigh(d_iu, self.symm)
        d_iu.reweigh(d_ui, self.symm)
        for d in (d_ui, d_iu):
            d.csr.refresh_weights_()
        for p in m.parameters():
            p.grad = None
        feats = {k: decoder.get_embed(m.embed_layers[k].weight, self.all_ids[k], self.noise[k], use_mask=True) for k in (u, i)}
        gt = {k: decoder.get_embed(m.embed_layers[k].weight, ids, None, use_mask=False) for k, ids in self.recon.items()}
        recon_rows = {u: self.graph[u, i].rows_of(self.recon[u]).contiguous() if u in self.recon else None,
                      i: self.graph[i, u].rows_of(self.recon[i]).contiguous() if i in self.recon else None}
        from .runtime import fork_join
        pred_ratings, pred_embeddings = [], []
        for b in range(m._n_blocks):
            layer = m.encoders[b][0]
            last = b == m._n_blocks - 1
            h, proj, pe, nxt = {}, {}, {}, {}

            def branch(k, other, d, rows):
                h[k] = layer.forward_single(k, feats[k], {other: (feats[other], d.csr, None, None, None)})
                head = m.rating_user_projs[b] if k == u else m.rating_item_projs[b]
                proj[k] = head(decoder.take_rows(h[k], rows))
                if k in self.recon:
                    pe[k] = m.embed_maps[b][k](h[k], recon_rows[k])
                if not last:
                    nxt[k] = m.embed_maps[b][k](h[k])

            with fork_join(self._streams) as run:
                run(0, lambda: branch(u, i, d_ui, rows_u))
                run(1, lambda: branch(i, u, d_iu, rows_i))
            pred_ratings.append(m.gen_ratings(proj[u], proj[i]))
            pred_embeddings.append({k: pe[k] for k in self.recon})
            if not last:
                feats = nxt
        loss = m.loss(pred_ratings, pred_embeddings, gt, self.ratings, self.mean, self.std, self.lam)
        loss.backward()
        self.loss = loss.detach()

    def capture(self):
        from .runtime import GraphedStep
        self._graph = GraphedStep(self._body)
        return self

    def load(self, pairs, ratings, noise=None, recon_ids=None):
        """Copy one iteration's inputs into the static buffers (asynchronous from pinned host memory)."""
        dev = self.dev
        pairs = pairs if isinstance(pairs, torch.Tensor) else torch.as_tensor(pairs)
        if tuple(pairs.shape) != (2, self.B):
            raise ValueError(f"pairs must have shape (2, {self.B}) — one captured graph per batch size")
        self.pairs.copy_(pairs.to(torch.int32), non_blocking=True)
        self.ratings.copy_(torch.as_tensor(ratings, dtype=torch.float32), non_blocking=True)
        for k, buf in self.noise.items():
            if noise is not None and k in noise:
                buf.copy_(torch.as_tensor(noise[k], dtype=torch.int32), non_blocking=True)
        for k, buf in self.recon.items():
            ids = _i32(recon_ids[k], dev) if not isinstance(recon_ids[k], torch.Tensor) else recon_ids[k]
            if ids.numel() != buf.numel():
                raise ValueError(f"recon ids of {k!r}: expected {buf.numel()}, got {ids.numel()}")
            buf.copy_(ids.to(torch.int32), non_blocking=True)

    def __call__(self, pairs, ratings, noise=None, recon_ids=None, eager=False):
        self.load(pairs, ratings, noise, recon_ids)
        if eager or self._graph is None:
            self._body()
        else:
            self._graph()
        return self.loss


__all__ = ["StaticGraphStep"]
