"""Node-partitioned multi-GPU aggregation: halo exchange of neighbour rows over NCCL / NVLink.

The reference has no multi-device path at all (single context, experiments/STAR-GCN.py:32;
SURVEY §2.2).  This module is the B200-native answer for one box of 8 GPUs:

  * both sides of the bipartite graph are split into contiguous id ranges, one per rank
    (``owner_ranges``); a rank owns the feature rows, the output rows and the CSR rows (all R
    rating levels) of its nodes
  * per layer direction the rows a rank's CSR references but does not own (the halo) are
    fetched with ONE variable-size all-to-all of D-float rows: owners pack the requested rows
    (deduplicated per peer — an item row referenced by 750 edges crosses NVLink once), receivers
    land them directly behind their local rows, and the fused gather kernel runs on the
    concatenated table ``[local rows ; halo rows]`` through column ids rewritten once per plan
  * backward is the transpose: halo-slot gradients go back through the reverse all-to-all and
    are summed into the owner's rows in a fixed order (sorted gather, no atomics)
  * when the halo is dense (at least half of all remote rows are needed by every rank — the case for a
    rating graph without locality) the exchange degenerates to an all-gather of the equal-sized blocks and
    its transpose to a reduce-scatter: ``HaloPlan(mode="auto")`` then uses those NVSwitch collectives
    directly, with no pack / unpack and the global ids as column ids
  * the small relation-weight gradients are all-reduced (inside the fused backward, or ``allreduce_grads``)

Index bookkeeping (``HaloPlan``) is host-side numpy + one all-to-all of index lists per plan and
runs under gloo on CPU tensors as well (tests/test_dist_cpu.py); the row traffic itself needs
the CUDA kernels.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import seg_op


def contiguous_ranges(n, world):
    """Equal contiguous id ranges [r_p, r_{p+1}) — (world+1,) int64."""
    return np.linspace(0, n, world + 1).round().astype(np.int64)


def balanced_ranges(degrees, world):
    """Contiguous ranges balanced by the number of edges (prefix sum of degrees / world)."""
    csum = np.concatenate([[0], np.cumsum(degrees, dtype=np.int64)])
    targets = csum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(csum, targets, side="left")
    return np.concatenate([[0], cuts, [len(degrees)]]).astype(np.int64)


def _all_to_all_ints(send_lists, group, device):
    """Variable-size all-to-all of int32 lists (index exchange, once per plan)."""
    world = dist.get_world_size(group)
    counts_out = torch.tensor([len(s) for s in send_lists], dtype=torch.int64, device=device)
    counts_in = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(counts_in, counts_out, group=group)
    in_splits = [int(c) for c in counts_in.cpu()]
    out_splits = [len(s) for s in send_lists]
    flat = np.concatenate(send_lists) if sum(out_splits) else np.zeros(0, np.int32)
    send = torch.from_numpy(flat.astype(np.int32)).to(device)
    recv = torch.empty(sum(in_splits), dtype=torch.int32, device=device)
    dist.all_to_all_single(recv, send, output_split_sizes=in_splits, input_split_sizes=out_splits, group=group)
    recv = recv.cpu().numpy()
    offs = np.concatenate([[0], np.cumsum(in_splits)])
    return [recv[offs[q]:offs[q + 1]] for q in range(world)]


class HaloPlan:
    """Who needs which rows of one node type's feature table, for one rank.

    cols_global   every column id (global) the rank's CSR references
    owner_ranges  (world+1,) contiguous ownership of the global ids
    Results:
      local_cols  (nnz,) int32  ids rewritten into [local rows ; halo rows] order
      recv_ids[q] sorted distinct global ids fetched from rank q   (recv_counts)
      send_idx[q] local row ids rank q fetches from this rank       (send_counts)
    ``exchange_fn(send_lists) -> recv_lists`` performs the index all-to-all; the default uses
    torch.distributed on ``index_device`` (cuda for NCCL, cpu for gloo).
    """

    def __init__(self, cols_global, owner_ranges, rank, world, group=None, index_device="cpu", exchange_fn=None,
                 mode="auto"):
        cols = np.asarray(cols_global, dtype=np.int64)
        self.rank, self.world = int(rank), int(world)
        self.owner_ranges = np.asarray(owner_ranges, dtype=np.int64)
        lo, hi = self.owner_ranges[rank], self.owner_ranges[rank + 1]
        self.n_local = int(hi - lo)
        self.group = group
        self._dev = None
        owner = np.searchsorted(self.owner_ranges, cols, side="right") - 1
        if cols.size and (cols.min() < 0 or cols.max() >= self.owner_ranges[-1]):
            raise ValueError("column id outside the partitioned id space")
        if mode not in ("auto", "alltoall", "allgather"):
            raise ValueError("mode must be 'auto', 'alltoall' or 'allgather'")
        self.mode = self._choose_mode(mode, cols, owner, index_device, exchange_fn)
        if self.mode == "allgather":
            # dense halo: every rank fetches (almost) every remote row, so the exchange is an all-gather of the
            # equal-sized blocks and its transpose a reduce-scatter (NVSwitch collectives, no pack / unpack);
            # the concatenated table is indexed by the global ids themselves
            self.local_cols = cols.astype(np.int32)
            self.n_ext = int(self.owner_ranges[-1])
            self.n_halo = self.n_ext - self.n_local
            self.recv_ids = [np.arange(self.owner_ranges[q], self.owner_ranges[q + 1]) if q != rank else np.zeros(0, np.int64)
                             for q in range(self.world)]
            self.recv_counts = [int(r.size) for r in self.recv_ids]
            self.send_idx = [np.arange(self.n_local, dtype=np.int32) if q != rank else np.zeros(0, np.int32)
                             for q in range(self.world)]
            self.send_counts = [int(s_.size) for s_ in self.send_idx]
            return
        self.recv_ids, local_cols = [], np.empty(cols.shape, np.int64)
        mine = owner == rank
        local_cols[mine] = cols[mine] - lo
        off = self.n_local
        for q in range(self.world):
            if q == rank:
                self.recv_ids.append(np.zeros(0, np.int64))
                continue
            sel = owner == q
            ids = np.unique(cols[sel])
            self.recv_ids.append(ids)
            local_cols[sel] = off + np.searchsorted(ids, cols[sel])
            off += ids.size
        self.n_halo = int(off - self.n_local)
        self.n_ext = int(off)
        self.local_cols = local_cols.astype(np.int32)
        self.recv_counts = [int(r.size) for r in self.recv_ids]
        # tell every owner which of ITS rows (local numbering there) this rank fetches
        requests = [(self.recv_ids[q] - self.owner_ranges[q]).astype(np.int32) for q in range(self.world)]
        if self.world == 1:
            got = [np.zeros(0, np.int32)]
        elif exchange_fn is not None:
            got = exchange_fn(requests)
        else:
            got = _all_to_all_ints(requests, group, torch.device(index_device))
        self.send_idx = [np.asarray(g, np.int32) for g in got]
        self.send_counts = [int(s.size) for s in self.send_idx]
        for q, s in enumerate(self.send_idx):
            if s.size and (s.min() < 0 or s.max() >= self.n_local):
                raise ValueError(f"rank {q} requested a row this rank does not own")

    def _choose_mode(self, mode, cols, owner, index_device, exchange_fn):
        """'allgather' needs equal blocks; 'auto' picks it when at least half of all remote rows are needed by
        EVERY rank (decided collectively so that all ranks take the same path)."""
        sizes = np.diff(self.owner_ranges)
        equal = bool(np.all(sizes == sizes[0]))
        if self.world == 1 or mode == "alltoall":
            return "alltoall"
        if mode == "allgather":
            if not equal:
                raise ValueError("allgather mode needs equal-sized ownership blocks")
            return "allgather"
        remote = owner != self.rank
        n_needed = np.unique(cols[remote]).size
        n_remote = int(self.owner_ranges[-1]) - self.n_local
        want = int(equal and n_remote > 0 and n_needed >= 0.5 * n_remote)
        if exchange_fn is not None or not dist.is_initialized():
            return "allgather" if want else "alltoall"
        flag = torch.tensor([want], dtype=torch.int32, device=torch.device(index_device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return "allgather" if int(flag.item()) else "alltoall"

    # ---- device-side state for the row exchange ----
    def to(self, device):
        cat = np.concatenate(self.send_idx) if sum(self.send_counts) else np.zeros(0, np.int32)
        self._dev = dict(send_cat=torch.from_numpy(cat).to(device), device=torch.device(device), pattern=None)
        return self

    def send_pattern(self):
        """CSR pattern with one edge per send slot, built once: forward = pack the requested rows,
        stable transpose = sum the halo gradients that come back."""
        d = self._dev
        if d["pattern"] is None:
            n = d["send_cat"].numel()
            d["pattern"] = seg_op.CSRPattern(d["send_cat"], torch.arange(n + 1, dtype=torch.int32, device=d["device"]),
                                             self.n_local, use_schedule=False)
            d["ones"] = torch.ones((1, n), dtype=torch.float32, device=d["device"])
        return d["pattern"]

    @property
    def halo_bytes_per_row_float(self):
        return 4 * self.n_halo


def _a2a_rows(out, inp, out_splits, in_splits, group):
    """Variable-size all-to-all of feature rows.  NCCL moves them GPU to GPU over NVLink; under a gloo
    group (the 1-GPU multi-process tests) the rows are staged through host memory."""
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
    else:
        host = torch.empty(out.shape, dtype=out.dtype)
        dist.all_to_all_single(host, inp.cpu(), output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
        out.copy_(host)


def _all_gather_rows(out, inp, group):
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, inp, group=group)
    else:  # gloo test transport
        world = dist.get_world_size(group)
        parts = [torch.empty(inp.shape, dtype=inp.dtype) for _ in range(world)]
        dist.all_gather(parts, inp.cpu(), group=group)
        out.copy_(torch.cat(parts))


def _reduce_scatter_rows(out, inp, group):
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(out, inp, group=group)
    else:  # gloo test transport
        host = inp.cpu()
        dist.all_reduce(host, group=group)
        r, n = dist.get_rank(group), out.shape[0]
        out.copy_(host[r * n:(r + 1) * n])


class _HaloExchange(torch.autograd.Function):
    """x_local [n_local, D] -> x_ext [n_local + n_halo, D] (local rows, then halo rows by owner, by id)."""

    @staticmethod
    def forward(ctx, x_local, plan):
        d = plan._dev
        D = x_local.shape[1]
        x_ext = torch.empty((plan.n_ext, D), dtype=torch.float32, device=x_local.device)
        ctx.plan, ctx.D = plan, D
        if plan.mode == "allgather":
            _all_gather_rows(x_ext, x_local, plan.group)
            return x_ext
        x_ext[:plan.n_local].copy_(x_local)
        if plan.world > 1:
            if d["send_cat"].numel():   # pack: one gather launch over the cached one-edge-per-slot pattern
                send = seg_op._seg_pool_fwd(x_local.unsqueeze(0), plan.send_pattern(), "sum")[0][0]
            else:
                send = torch.empty((0, D), dtype=torch.float32, device=x_local.device)
            _a2a_rows(x_ext[plan.n_local:], send, plan.recv_counts, plan.send_counts, plan.group)
        return x_ext

    @staticmethod
    def backward(ctx, g_ext):
        plan, D = ctx.plan, ctx.D
        if plan.mode == "allgather":
            g_local = torch.empty((plan.n_local, D), dtype=torch.float32, device=g_ext.device)
            _reduce_scatter_rows(g_local, g_ext.contiguous(), plan.group)
            return g_local, None
        g_local = g_ext[:plan.n_local].clone()
        if plan.world > 1:
            n_send = sum(plan.send_counts)
            g_back = torch.empty((n_send, D), dtype=torch.float32, device=g_ext.device)
            _a2a_rows(g_back, g_ext[plan.n_local:].contiguous(), plan.send_counts, plan.recv_counts, plan.group)
            if n_send:
                pat = plan.send_pattern()
                seg_op._weighted_pool_bwd_data(g_back.unsqueeze(0), plan._dev["ones"], pat, plan.n_local,
                                               out=g_local.unsqueeze(0), req="add")
        return g_local, None


def halo_exchange(x_local, plan):
    """Rows of the neighbour table this rank's CSR needs: ``[x_local ; rows fetched from peers]``."""
    if plan._dev is None:
        plan.to(x_local.device)
    if x_local.shape[0] != plan.n_local:
        raise ValueError(f"x_local has {x_local.shape[0]} rows, the plan owns {plan.n_local}")
    return _HaloExchange.apply(x_local.contiguous(), plan)


def allreduce_grads(params, group=None):
    """Sum the (small) parameter gradients over ranks through one flat buffer."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, group=group)
    else:
        host = flat.cpu()
        dist.all_reduce(host, group=group)
        flat.copy_(host)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


# ------------------------------------------------------------------------------------------------
# Synthetic partitioned workload: every rank holds an ML-10M-shaped slice of a world-times larger graph
# ------------------------------------------------------------------------------------------------
def edge_owner(u, i, p, world):
    """Deterministic pseudo-random owner of the item end of base edge (u, i) as seen from user block p."""
    h = (u.astype(np.int64) * 2654435761 + i.astype(np.int64) * 40503 + p * 97) >> 7
    return (h % world).astype(np.int64)


def partitioned_layer_inputs(base, rank, world):
    """Global graph = ``world`` user blocks x ``world`` item blocks built from one base bipartite graph
    (stargcn_b200.synth.make_bipartite): base edge (u, i) in user block p points at item i of block
    ``edge_owner(u, i, p)``.  Every rank can derive its own rows of BOTH directions from the base graph
    alone.  Returns {'user': (indptr, cols_global, vals, support), 'item': (...)} for this rank, with
    global column ids, plus the ownership ranges."""
    nu, ni = base["n_user"], base["n_item"]
    u2i, i2u = base["u2i"], base["i2u"]
    out = dict(user_ranges=np.arange(world + 1, dtype=np.int64) * nu, item_ranges=np.arange(world + 1, dtype=np.int64) * ni)
    # user rows of block `rank`: same pattern as the base graph, item ends scattered over the item blocks
    rows = u2i["rows"].astype(np.int64)
    cols = u2i["cols"].astype(np.int64)
    q = edge_owner(rows, cols, rank, world)
    gcols = q * ni + cols
    order = np.lexsort((gcols, rows))  # column ids sorted inside each row, as scipy tocsr gives them
    out["user"] = (u2i["indptr"], gcols[order], u2i["vals"][order], u2i["support"][order])
    # item rows of block `rank`: base edge (u, i) of user block p lands here when edge_owner == rank
    r_l, c_l, v_l, s_l = [], [], [], []
    irows, icols = i2u["rows"].astype(np.int64), i2u["cols"].astype(np.int64)
    for p in range(world):
        sel = edge_owner(icols, irows, p, world) == rank
        r_l.append(irows[sel]); c_l.append(p * nu + icols[sel]); v_l.append(i2u["vals"][sel]); s_l.append(i2u["support"][sel])
    r, c = np.concatenate(r_l), np.concatenate(c_l)
    v, s = np.concatenate(v_l), np.concatenate(s_l)
    order = np.lexsort((c, r))
    r, c, v, s = r[order], c[order], v[order], s[order]
    indptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=ni))]).astype(np.int32)
    out["item"] = (indptr, c, v, s)
    return out


__all__ = ["HaloPlan", "halo_exchange", "allreduce_grads", "contiguous_ranges", "balanced_ranges",
           "partitioned_layer_inputs", "edge_owner"]
