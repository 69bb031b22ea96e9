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
  * ``mode="peer"`` (picked by ``"auto"`` for a dense halo when symmetric memory is available) replaces the NCCL
    collectives of the dense case by this library's own kernels over NVLink peer memory (csrc/peer.cu,
    ``PeerTransport``): all-gather = every rank stores its block into every rank's table, reduce-scatter = the
    transposed gather stores each gradient row straight into its owner's staging slot (the transfer rides inside
    the gather launch) + a local fixed-order sum, gradient all-reduce = push + the same sum, one flag barrier each

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
        if mode not in ("auto", "alltoall", "allgather", "peer", "peer_dense", "peer_sparse", "nccl"):
            raise ValueError("mode must be 'auto', 'nccl', 'alltoall', 'allgather', 'peer', 'peer_dense' or 'peer_sparse'")
        self._transport = None
        # mode: the transport ('alltoall' / 'allgather' over NCCL, 'peer' = this library's kernels over NVLink peer
        # memory); dense: the table layout — every row of every rank under its global id, or [local rows ; the
        # deduplicated halo rows by owner, by id]
        self.mode, self.dense = self._choose_mode(mode, cols, owner, index_device, exchange_fn)
        if self.dense:
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
        """-> (transport, dense layout?).  'allgather' needs equal blocks; 'peer' (this library's kernels over NVLink
        peer memory) takes any contiguous ranges and both layouts.  A halo is DENSE when at least half of all remote
        rows are needed by EVERY rank (decided collectively so that all ranks take the same path): 'auto' then uses
        the global-id table — over peer memory when the ranks are CUDA devices of one box (at most SG_MAX_PEERS), else
        the NCCL all-gather for equal blocks — and the deduplicated halo rows otherwise (peer memory, else the NCCL
        all-to-all).  'nccl' = 'auto' without the peer transport; 'peer_dense' / 'peer_sparse' force a layout."""
        sizes = np.diff(self.owner_ranges)
        equal = bool(np.all(sizes == sizes[0]))
        if mode in ("peer", "peer_dense", "peer_sparse") and self.world > MAX_PEERS:
            raise ValueError(f"peer mode handles at most {MAX_PEERS} ranks (one NVSwitch box)")
        if mode == "peer_dense":
            return "peer", True
        if mode == "peer_sparse":
            return "peer", False
        if self.world == 1 or mode == "alltoall":
            return "alltoall", False
        if mode == "allgather":
            if not equal:
                raise ValueError("allgather mode needs equal-sized ownership blocks")
            return "allgather", True
        remote = owner != self.rank
        n_needed = np.unique(cols[remote]).size
        n_remote = int(self.owner_ranges[-1]) - self.n_local
        dense = int(n_remote > 0 and n_needed >= 0.5 * n_remote)
        collective = exchange_fn is None and dist.is_initialized()
        if collective:
            flag = torch.tensor([dense], dtype=torch.int32, device=torch.device(index_device))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            dense = int(flag.item())
        peer_ok = mode == "peer" or (mode == "auto" and collective and self.world <= MAX_PEERS
                                     and torch.device(index_device).type == "cuda" and dist.get_backend(self.group) == "nccl")
        if peer_ok:
            return "peer", bool(dense)
        if dense and equal:
            return "allgather", True
        return "alltoall", False

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

    def try_peer_transport(self, D, grad_floats, device):
        """Create the peer transport NOW and agree collectively that every rank succeeded.  Returns True, or False
        after resetting nothing — the caller then rebuilds its plans with ``mode='nccl'`` (symmetric memory needs
        peer-to-peer capable CUDA devices of one box; anything else falls back to the NCCL collectives)."""
        ok = 1
        try:
            self.peer_transport(D, grad_floats, device)
        except Exception as e:                      # noqa: BLE001 — any failure means "no peer memory here"
            import sys
            print(f"[stargcn_b200.dist] rank {self.rank}: peer-memory transport unavailable ({type(e).__name__}: {e}); "
                  f"falling back to NCCL", file=sys.stderr)
            self._transport, ok = None, 0
        if dist.is_initialized() and self.world > 1:
            flag = torch.tensor([ok], dtype=torch.int32, device=torch.device(device))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            ok = int(flag.item())
        if not ok:
            self._transport = None
        return bool(ok)

    def peer_transport(self, D, grad_floats, device):
        """The symmetric-memory buffers of this direction (mode 'peer'), created collectively on first use."""
        t = self._transport
        if t is None or t.D != D or t.grad_capacity < grad_floats:
            t = self._transport = PeerTransport(self, D, grad_floats, device)
        return t


MAX_PEERS = 8          # SG_MAX_PEERS (include/stargcn_b200.h)

# None (shipped): the all-gather is the store kernel (sg_peer_push_rows).  A byte count: blocks of at least that size go
# to the peers by the COPY ENGINES instead (one asynchronous device-to-device copy per peer into its mapped table; no SM
# is used).  Measured on 8 B200s (weak scaling, 17.9 MB block to 8 tables): store kernel 0.205 ms / step 1.870 ms, copy
# engines 0.235 ms / step 1.954 ms — seven serial copies do not fill the NVLink ports the way 296 storing CTAs do
# (profiles/r02_summary.md §G); at 2 GPUs both give 1.46 ms.  Kept as a tested option (bench.py --peer-push ce).
PEER_COPY_ENGINE_BYTES = None


def peer_sparse_layout(counts, rank):
    """Row arithmetic of the sparse-halo peer exchange for one rank (pure; tests/test_dist_cpu.py).

    counts[p] = dict(n_local, recv=[rows p fetches from q], send=[rows p sends to q]) of EVERY rank p, with
    recv[p][q] == send[q][p].  A rank's table is [own rows ; rows fetched from rank 0, 1, ... (by id)] and its send
    slots are ordered by requesting rank.  Returns, in ROWS:
      send_lo    (W+1) slot ranges of the send list by requesting rank p
      x_dst[p]   first row of p's table that this rank's block lands in (forward pack-and-push)
      halo_lo    (W+2) row ranges of this rank's table by gradient target: [own rows | fetched from 0 | from 1 | ...]
      g_dst[q]   first row of q's gradient staging that the slots fetched from q go back to (= where q's send list
                 holds the block it sent to this rank)
      n_ext, n_send, x_rows (largest table over ranks), g_rows (largest send list over ranks)"""
    W = len(counts)
    me = counts[rank]
    send_lo = [0]
    for p_ in range(W):
        send_lo.append(send_lo[-1] + int(me["send"][p_]))
    halo_lo = [0, int(me["n_local"])]
    for q in range(W):
        halo_lo.append(halo_lo[-1] + int(me["recv"][q]))
    return dict(send_lo=send_lo, halo_lo=halo_lo,
                x_dst=[int(counts[p_]["n_local"]) + sum(int(c) for c in counts[p_]["recv"][:rank]) for p_ in range(W)],
                g_dst=[sum(int(c) for c in counts[q]["send"][:rank]) for q in range(W)],
                n_ext=halo_lo[-1], n_send=send_lo[-1],
                x_rows=max(int(c["n_local"]) + sum(int(v) for v in c["recv"]) for c in counts),
                g_rows=max(sum(int(v) for v in c["send"]) for c in counts))


class PeerTransport:
    """Exchange buffers of ONE layer direction in symmetric memory: every rank maps every rank's buffer, so the
    collectives of the partitioned step are this library's own kernels over NVLink peer memory (csrc/peer.cu).

    Layout of each rank's buffer (fp32 words): [64 flag words | x_ext | g_stage | w_stage [world, grad_capacity]].
      dense halo   x_ext [n_total, D] under global ids: rank p stores block p into every rank's table (all-gather);
                   g_stage [world, slot]: the transposed gather of rank p stores row j into slot p of j's owner and the
                   owner sums its slots in rank order (reduce-scatter)
      sparse halo  x_ext [n_local + n_halo, D]: the rank's own rows, then the DEDUPLICATED rows it needs from every
                   peer (by owner, by id); one gather launch packs the rows each peer asked for and stores them
                   straight into that peer's halo slots (all-to-all); g_stage [n_send, D]: the transposed gather stores
                   the gradient of every halo slot straight into its owner's g_stage at the position of the matching
                   send slot, and the owner adds them to its rows by the sorted-transpose gather (fixed order)
    The weight gradient goes through w_stage (push to every rank's slot, sum in rank order).  One stream per
    direction; see peer.cu for why single buffers suffice."""
    FLAG_WORDS = 64

    def __init__(self, plan, D, grad_floats, device, timeout_s=10.0):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self._lib, self._ctypes = _lib, ctypes
        self.plan, self.D, self.timeout_s = plan, int(D), float(timeout_s)
        self.rank, self.world, self.dense = plan.rank, plan.world, bool(plan.dense)
        self.device = torch.device(device)
        if self.world > MAX_PEERS:
            raise ValueError(f"peer transport handles at most {MAX_PEERS} ranks")
        W, rank, D = self.world, self.rank, self.D
        lo = [int(v) for v in plan.owner_ranges]
        self.n_total, self.n_local = lo[-1], lo[rank + 1] - lo[rank]
        group = plan.group if plan.group is not None else dist.group.WORLD

        def al(n):
            return (int(n) + 63) // 64 * 64

        if self.dense:
            x_rows, g_floats = self.n_total, None
            self.slot = al(max(b - a for a, b in zip(lo[:-1], lo[1:])) * D)
            g_floats = W * self.slot
        else:
            # every rank's counts: recv[p][q] = rows p fetches from q (= rows q sends to p)
            mine = dict(n_local=self.n_local, recv=[int(c) for c in plan.recv_counts], send=[int(c) for c in plan.send_counts])
            allc = [None] * W
            dist.all_gather_object(allc, mine, group=group)
            for p_ in range(W):
                for q in range(W):
                    if allc[p_]["recv"][q] != allc[q]["send"][p_]:
                        raise RuntimeError("halo plans of the ranks do not match (recv / send counts differ)")
            self._counts = allc
            self._lay = lay = peer_sparse_layout(allc, rank)
            x_rows = lay["x_rows"]                      # symmetric allocation: the largest rank's size
            g_floats = al(lay["g_rows"] * D)
            self.n_ext, self.n_send = lay["n_ext"], lay["n_send"]
        self.grad_capacity = al(grad_floats)
        off, pos = {}, self.FLAG_WORDS
        for name, n in (("x_ext", al(x_rows * D)), ("g_stage", g_floats), ("w_stage", W * self.grad_capacity)):
            off[name], pos = pos, pos + n
        self.off = off
        self.buf = symm.empty(pos, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        base = [int(b) for b in self.handle.buffer_ptrs]
        if len(base) != W or base[rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned an unexpected pointer table")
        self._base = base
        self.state = torch.zeros(2, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)            # every rank's flag words are zero before anyone can arrive

        def table(vals):
            return (ctypes.c_void_p * len(vals))(*vals)

        self._table = table
        self._flags = table(base)
        self._w_dst = table([b + 4 * (off["w_stage"] + rank * self.grad_capacity) for b in base])
        self._w_stage = self.buf[off["w_stage"]:off["w_stage"] + W * self.grad_capacity]
        if self.dense:
            self._x_dst = table([b + 4 * (off["x_ext"] + lo[rank] * D) for b in base])
            self._g_dst = table([b + 4 * (off["g_stage"] + rank * self.slot) for b in base])
            self._owner_lo = (ctypes.c_int32 * (W + 1))(*lo)
            self.x_ext = self.buf[off["x_ext"]:off["x_ext"] + self.n_total * D].view(self.n_total, D)
            self._g_stage = self.buf[off["g_stage"]:off["g_stage"] + W * self.slot]
            # this rank's block inside every rank's table, as tensors (copy-engine path of all_gather)
            self._x_peer = [self.handle.get_buffer(q, (self.n_local, D), torch.float32, off["x_ext"] + lo[rank] * D)
                            for q in range(W)]
        else:
            self.x_ext = self.buf[off["x_ext"]:off["x_ext"] + self.n_ext * D].view(self.n_ext, D)
            self._g_stage = self.buf[off["g_stage"]:off["g_stage"] + max(self.n_send, 1) * D].view(max(self.n_send, 1), D)[:self.n_send]
            # forward: the block of send slots for peer p lands in p's table behind p's own rows and the rows p fetches
            # from lower ranks; backward: rows [0, n_local) stay here (target 0, set per call), the halo slots fetched
            # from owner q go to q's g_stage at the position of the block q sends to this rank (peer_sparse_layout)
            lay = self._lay
            self._send_lo = (ctypes.c_int32 * (W + 1))(*lay["send_lo"])
            self._x_dst = table([base[p_] + 4 * (off["x_ext"] + lay["x_dst"][p_] * D) for p_ in range(W)])
            self._halo_lo = (ctypes.c_int32 * (W + 2))(*lay["halo_lo"])
            self._g_dst_peers = [base[q] + 4 * (off["g_stage"] + lay["g_dst"][q] * D) for q in range(W)]
            d = plan._dev if plan._dev is not None else plan.to(self.device)._dev
            self._send_cat = d["send_cat"]
            self._send_ptr = torch.arange(self.n_send + 1, dtype=torch.int32, device=self.device)
            self._send_ones = torch.ones(max(self.n_send, 1), dtype=torch.float32, device=self.device)
        self._open = False                   # a forward whose consumers no barrier has covered yet
        self._g_self = None

    def _prof(self, tag, fn):
        from . import graph
        e0 = graph._prof_begin()
        fn()
        graph._prof_end(tag, e0, self.plan)

    def _stream(self):
        return self._ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _ptr(self, t):
        return self._ctypes.c_void_p(t.data_ptr())

    def barrier(self):
        lib = self._lib.load()
        self._prof("peer_barrier", lambda: self._lib.check(
            lib.sg_peer_barrier(self._flags, self._ptr(self.state), self.rank, self.world, self.timeout_s, self._stream()),
            "sg_peer_barrier"))

    def all_gather(self, x_local, will_backward):
        """x_local [n_local, D] -> this rank's table x_ext (a view of the symmetric buffer, valid until the next
        call): [n_total, D] under global ids (dense) or [own rows ; deduplicated halo rows] (sparse).
        ``will_backward``: the backward's barrier will cover this table's readers; otherwise the caller ends its
        forward with :meth:`release`."""
        if x_local.shape != (self.n_local, self.D) or x_local.dtype != torch.float32 or not x_local.is_contiguous():
            raise ValueError(f"x_local must be a contiguous float32 [{self.n_local}, {self.D}] tensor")
        if self._open:                       # the previous forward never reached a covering barrier
            self.barrier()
        lib, c = self._lib.load(), self._ctypes
        n = self.n_local * self.D

        def push_dense():
            if PEER_COPY_ENGINE_BYTES is not None and 4 * n >= PEER_COPY_ENGINE_BYTES:
                for k in range(self.world):          # start with the next rank so that the links are used evenly
                    self._x_peer[(self.rank + 1 + k) % self.world].copy_(x_local, non_blocking=True)
            else:
                self._lib.check(lib.sg_peer_push_rows(self._x_dst, self._ptr(x_local), n, self.world, self._stream()),
                                "sg_peer_push_rows")

        def push_sparse():
            # own rows: a local copy; requested rows: ONE gather launch (one edge per send slot, unit weights) whose
            # output rows are stored straight into the requesting ranks' halo slots
            own = self._table([self.x_ext.data_ptr()])
            self._lib.check(lib.sg_peer_push_rows(own, self._ptr(x_local), n, 1, self._stream()), "sg_peer_push_rows")
            if self.n_send:
                self._lib.check(lib.sg_multilink_agg_bwd_peer(
                    self._x_dst, self._send_lo, self.world, self._ptr(x_local), self._ptr(self._send_ones),
                    self._ptr(self._send_cat), self._ptr(self._send_ptr), 1, self.n_local, self.n_send, self.n_send, self.D,
                    c.c_void_p(0), 0, c.c_void_p(0), self._stream()), "sg_multilink_agg_bwd_peer (pack + push)")

        self._prof("peer_push", push_dense if self.dense else push_sparse)
        self.barrier()
        self._open = True
        return self.x_ext

    def release(self):
        """Forward-only use: the readers of x_ext have been issued; no rank may overwrite it before all arrive here."""
        self.barrier()
        self._open = False

    def push_grad(self, g_flat):
        """Store this rank's packed weight gradient (flat, <= grad_capacity floats, padded to 4) into every rank's slot."""
        n = (g_flat.numel() + 3) // 4 * 4
        if n > self.grad_capacity or g_flat.untyped_storage().nbytes() - 4 * g_flat.storage_offset() < 4 * n:
            raise ValueError("gradient buffer larger than the staging slot or not padded to 4 floats")
        lib = self._lib.load()
        self._lib.check(lib.sg_peer_push_rows(self._w_dst, self._ptr(g_flat), n, self.world, self._stream()),
                        "sg_peer_push_rows")

    def scatter_args(self):
        """(target pointer table, row ranges, number of targets) for sg_multilink_agg_bwd_peer — where the transposed
        gather stores the gradient row of every table row."""
        if self.dense:
            return self._g_dst, self._owner_lo, self.world
        # target 0: this rank's own rows, written locally into the tensor reduce_rows() will return
        self._g_self = torch.empty((self.n_local, self.D), dtype=torch.float32, device=self.device)
        targets = [self._g_self.data_ptr() if self.n_local else self.x_ext.data_ptr()] + self._g_dst_peers
        self._targets = self._table(targets)         # kept alive until the launch has been issued
        return self._targets, self._halo_lo, self.world + 1

    def reduce_rows(self):
        """Gradient of this rank's own rows after the backward barrier -> [n_local, D].  Dense: the world staging
        slots summed in rank order.  Sparse: the locally written rows + the halo gradients the peers stored into
        g_stage, added by the sorted-transpose gather over the send pattern (fixed order)."""
        lib = self._lib.load()
        if self.dense:
            out = torch.empty((self.n_local, self.D), dtype=torch.float32, device=self.device)
            if self.n_local:
                n = self.n_local * self.D
                self._prof("peer_reduce", lambda: self._lib.check(
                    lib.sg_peer_reduce(self._ptr(out), self._ptr(self._g_stage), n, self.slot, self.world, 1, self._stream()),
                    "sg_peer_reduce"))
            return out
        out, self._g_self = self._g_self, None
        if self.n_send and self.n_local:
            pat = self.plan.send_pattern()
            self._prof("peer_reduce", lambda: seg_op._weighted_pool_bwd_data(
                self._g_stage.unsqueeze(0), self.plan._dev["ones"], pat, self.n_local, out=out.unsqueeze(0), req="add"))
        return out

    def reduce_grad(self, g_flat):
        """g_flat <- sum over ranks of the pushed gradients (rank order: the same bits on every rank)."""
        n = (g_flat.numel() + 3) // 4 * 4
        lib = self._lib.load()
        self._lib.check(lib.sg_peer_reduce(self._ptr(g_flat), self._ptr(self._w_stage), n, self.grad_capacity, self.world, 1,
                                           self._stream()), "sg_peer_reduce")
        return g_flat

    def check(self):
        """Raise if a barrier timed out (synchronises)."""
        err = int(self.state[1].item())
        if err:
            raise RuntimeError(f"peer barrier timed out waiting for rank {err - 1} (direction group of rank {self.rank})")


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
        if plan.mode == "peer":
            raise RuntimeError("mode 'peer' is driven by the fused aggregation (MultiLinkGCNAggregator.halo_plan), "
                               "not by halo_exchange()")
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


def partitioned_aggregate(agg, x_local, plan, *plan_lists):
    """One partitioned layer direction: ``agg`` (a MultiLinkGCNAggregator) over the rank's own neighbour rows
    ``x_local`` and its CSR (a MultiLinkCSR or the three per-level lists, column ids = ``plan.local_cols``).
    Mode 'peer' runs the exchange inside the fused op (NVLink peer memory); the NCCL modes exchange first."""
    if plan is None:
        agg.halo_plan = None
        return agg(x_local, *plan_lists)
    if plan.mode == "peer":
        agg.halo_plan = plan
        return agg(x_local, *plan_lists)
    agg.halo_plan = None
    return agg(halo_exchange(x_local, plan), *plan_lists)


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


__all__ = ["HaloPlan", "PeerTransport", "peer_sparse_layout", "halo_exchange", "partitioned_aggregate", "allreduce_grads", "contiguous_ranges", "balanced_ranges",
           "partitioned_layer_inputs", "edge_owner"]
