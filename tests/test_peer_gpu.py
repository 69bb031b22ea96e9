"""The NVLink peer-memory exchange kernels (csrc/peer.cu + the PEER epilogue of the gather) on ONE device: the `world`
ranks are simulated by `world` buffers and `world` streams of one process, so every kernel runs exactly as it does
between processes (the pointer tables simply address local memory).  The real multi-process path over NVLink is
checked by ``bench.py --gpus N --check`` (tests/test_dist_gpu.py, needs >= 2 GPUs)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from stargcn_b200 import _lib
    return _lib, _lib.load()


def _table(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


def _s(stream):
    return ctypes.c_void_p(stream.cuda_stream)


@pytest.mark.parametrize("world", [1, 2, 5, 8])
def test_push_barrier_reduce(world):
    L, lib = _lib()
    dev = torch.device("cuda")
    D = 64
    rs = np.random.RandomState(world)
    sizes = rs.randint(3, 40, size=world)            # unequal blocks
    lo = np.concatenate([[0], np.cumsum(sizes)])
    n_tot = int(lo[-1])
    blocks = [torch.from_numpy(rs.normal(size=(int(n), D)).astype(np.float32)).to(dev) for n in sizes]
    tables = [torch.full((n_tot, D), float("nan"), device=dev) for _ in range(world)]
    flags = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(world)]
    states = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    flag_tab = _table([f.data_ptr() for f in flags])
    torch.cuda.synchronize()
    for step in range(3):                             # epochs advance; buffers reused
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                dst = _table([t.data_ptr() + 4 * int(lo[r]) * D for t in tables])
                L.check(lib.sg_peer_push_rows(dst, ctypes.c_void_p(blocks[r].data_ptr()), int(sizes[r]) * D, world, _s(streams[r])),
                        "push")
                L.check(lib.sg_peer_barrier(flag_tab, ctypes.c_void_p(states[r].data_ptr()), r, world, 5.0, _s(streams[r])), "barrier")
        torch.cuda.synchronize()
        want = torch.cat(blocks)
        for r in range(world):
            assert torch.equal(tables[r], want)
            assert states[r].tolist() == [step + 1, 0]
    # local half of the reduce-scatter: slots summed in rank order, bit-exact against the same additions in torch
    n, slot = 37 * D, 40 * D
    stage = torch.from_numpy(rs.normal(size=(world, slot)).astype(np.float32)).to(dev)
    out = torch.empty(n, device=dev)
    L.check(lib.sg_peer_reduce(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stage.data_ptr()), n, slot, world, 1, None), "reduce")
    ref = stage[0, :n].clone()
    for q in range(1, world):
        ref = ref + stage[q, :n]
    assert torch.equal(out, ref)
    base = torch.from_numpy(rs.normal(size=n).astype(np.float32)).to(dev)
    acc = base.clone()
    L.check(lib.sg_peer_reduce(ctypes.c_void_p(acc.data_ptr()), ctypes.c_void_p(stage.data_ptr()), n, slot, world, 3, None), "reduce add")
    assert torch.equal(acc, ref + base)


def test_barrier_timeout_reports_missing_rank():
    L, lib = _lib()
    dev = torch.device("cuda")
    flags = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(3)]
    state = torch.zeros(2, dtype=torch.int32, device=dev)
    tab = _table([f.data_ptr() for f in flags])
    L.check(lib.sg_peer_barrier(tab, ctypes.c_void_p(state.data_ptr()), 0, 3, 0.05, None), "barrier")   # ranks 1, 2 never arrive
    torch.cuda.synchronize()
    assert state[1].item() in (2, 3) and state[0].item() == 1
    assert flags[1][0].item() == 1 and flags[2][0].item() == 1     # this rank's arrival reached both peers
    with pytest.raises(ValueError):
        L.check(lib.sg_peer_barrier(tab, ctypes.c_void_p(state.data_ptr()), 3, 3, 1.0, None), "barrier")
    with pytest.raises(ValueError):
        L.check(lib.sg_peer_push_rows(tab, ctypes.c_void_p(state.data_ptr()), 6, 3, None), "push")      # not a multiple of 4


@pytest.mark.parametrize("D", [64, 32])
def test_transposed_gather_scatters_rows_to_owner_slots(D):
    """sg_multilink_agg_bwd_peer == sg_multilink_agg_bwd, with row j landing in owner(j)'s staging buffer (bit-exact:
    the same kernel arithmetic, only the store address differs) — with and without a schedule (split segments)."""
    L, lib = _lib()
    from stargcn_b200 import synth
    from stargcn_b200.graph import MultiLinkCSR
    dev = torch.device("cuda")
    R = 5
    base = synth.make_bipartite(400, 90, 30_000, n_levels=R, seed=9)
    c = base["u2i"]
    lists = synth.split_by_level(c["indptr"], c["cols"], c["vals"], c["support"], base["levels"])[:3]
    n_nb = base["n_item"]
    world = 3
    owner_lo = np.array([0, 17, 60, n_nb], dtype=np.int32)
    for chunk in (256, 16):                          # 16: the hot items are cut into many partials -> combine pass
        csr = MultiLinkCSR(*lists, n_nb=n_nb, device=dev, chunk=chunk).prepare(backward=True)
        gagg = torch.randn((csr.n_dst, R * D), device=dev)
        t_indptr, t_src, t_w = csr.transposed()
        sched = csr.t_schedule()
        part = sched.partial(1, D)
        args = (ctypes.c_void_p(gagg.data_ptr()), ctypes.c_void_p(t_w.data_ptr()), ctypes.c_void_p(t_src.data_ptr()),
                ctypes.c_void_p(t_indptr.data_ptr()), R, csr.n_dst, n_nb, csr.nnz, D)
        tail = (ctypes.c_void_p(sched.buf.data_ptr()), sched.chunk, ctypes.c_void_p(part.data_ptr()), None)
        gx = torch.empty((n_nb, D), device=dev)
        L.check(lib.sg_multilink_agg_bwd(ctypes.c_void_p(gx.data_ptr()), *args, 1, *tail), "agg_bwd")
        stages = [torch.full((int(owner_lo[q + 1] - owner_lo[q]), D), float("nan"), device=dev) for q in range(world)]
        L.check(lib.sg_multilink_agg_bwd_peer(_table([s.data_ptr() for s in stages]), (ctypes.c_int32 * (world + 1))(*owner_lo.tolist()),
                                              world, *args, *tail), "agg_bwd_peer")
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(stages), gx)
    with pytest.raises(ValueError):                   # ranges must cover [0, n_nb)
        L.check(lib.sg_multilink_agg_bwd_peer(_table([s.data_ptr() for s in stages]), (ctypes.c_int32 * (world + 1))(0, 17, 60, n_nb - 1),
                                              world, *args, *tail), "agg_bwd_peer")


def test_pack_and_push_rows_with_empty_and_nine_targets():
    """The sparse-halo forward: one gather launch (one edge per send slot, unit weights, R = 1, no schedule) packs the
    requested rows and stores each block into its target's buffer — SG_MAX_PEERS + 1 targets, some of them empty."""
    L, lib = _lib()
    dev = torch.device("cuda")
    D, n_local = 64, 500
    rs = np.random.RandomState(4)
    x = torch.from_numpy(rs.normal(size=(n_local, D)).astype(np.float32)).to(dev)
    counts = [7, 0, 33, 0, 0, 120, 1, 64, 12]                       # 9 targets
    lo = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n_send = int(lo[-1])
    send = torch.from_numpy(rs.randint(0, n_local, n_send).astype(np.int32)).to(dev)
    ptr = torch.arange(n_send + 1, dtype=torch.int32, device=dev)
    ones = torch.ones(n_send, device=dev)
    bufs = [torch.full((max(c, 1), D), float("nan"), device=dev) for c in counts]
    L.check(lib.sg_multilink_agg_bwd_peer(_table([b.data_ptr() for b in bufs]), (ctypes.c_int32 * 10)(*lo.tolist()), 9,
                                          ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(ones.data_ptr()), ctypes.c_void_p(send.data_ptr()),
                                          ctypes.c_void_p(ptr.data_ptr()), 1, n_local, n_send, n_send, D, None, 0, None, None), "pack+push")
    torch.cuda.synchronize()
    for q, c in enumerate(counts):
        if c:
            assert torch.equal(bufs[q][:c], x[send[lo[q]:lo[q + 1]].long()])
        else:
            assert torch.isnan(bufs[q]).all()                        # an empty range receives nothing
    with pytest.raises(ValueError):                                  # more than SG_MAX_PEERS + 1 targets
        L.check(lib.sg_multilink_agg_bwd_peer(_table([b.data_ptr() for b in bufs] + [bufs[0].data_ptr()]),
                                              (ctypes.c_int32 * 11)(*lo.tolist(), n_send), 10,
                                              ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(ones.data_ptr()), ctypes.c_void_p(send.data_ptr()),
                                              ctypes.c_void_p(ptr.data_ptr()), 1, n_local, n_send, n_send, D, None, 0, None, None), "pack+push")


def _single_rank_worker(port, q):
    """world = 1 process group: the PeerTransport class end to end on one device (symmetric allocation, pointer tables,
    push / barrier / scatter / reduce launches inside the fused op, forward-only release) against the plain layer."""
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        from stargcn_b200 import dist as sgd, synth
        from stargcn_b200.graph import MultiLinkCSR
        from stargcn_b200.layers import MultiLinkGCNAggregator
        dev = torch.device("cuda", 0)
        R, D, U = 5, 64, 250
        base = synth.make_bipartite(300, 200, 6000, n_levels=R, seed=3)
        c = base["u2i"]
        res = {}
        for mode in ("peer_dense", "peer_sparse"):
            plan = sgd.HaloPlan(c["cols"], np.array([0, base["n_item"]]), 0, 1, index_device=dev, mode=mode).to(dev)
            assert plan.mode == "peer" and plan.n_ext == base["n_item"] and plan.n_halo == 0
            lists = synth.split_by_level(c["indptr"], plan.local_cols, c["vals"], c["support"], base["levels"])[:3]
            csr = MultiLinkCSR(*lists, n_nb=plan.n_ext, device=dev)
            torch.manual_seed(1)
            agg = MultiLinkGCNAggregator(units=U, num_links=R, act="leaky", accum="sum", in_units=D).to(dev)
            x = torch.randn((base["n_item"], D), device=dev, requires_grad=True)
            gout = torch.randn((base["n_user"], U), device=dev)
            ref = agg(x, csr)
            ref.backward(gout)
            want = (ref.detach().clone(), x.grad.clone(), agg.weight1.grad.clone())
            x.grad = None
            agg.zero_grad(set_to_none=True)
            agg.grad_group = dist.group.WORLD
            with torch.no_grad():
                out0 = sgd.partitioned_aggregate(agg, x, plan, csr)            # forward-only: ends with release()
            for _ in range(2):                                                   # buffers reused
                x.grad = None
                agg.zero_grad(set_to_none=True)
                out = sgd.partitioned_aggregate(agg, x, plan, csr)
                out.backward(gout)
            plan._transport.check()
            res[mode] = bool(torch.equal(out0, want[0]) and torch.equal(out, want[0]) and torch.equal(x.grad, want[1])
                             and torch.equal(agg.weight1.grad, want[2]))
        q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_peer_transport_single_rank_equals_plain_layer():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_single_rank_worker, args=(port, q))
    p.start()
    res = q.get(timeout=240)
    p.join(timeout=60)
    assert p.exitcode == 0
    assert res == {"peer_dense": True, "peer_sparse": True}
