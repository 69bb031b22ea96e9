"""Partitioned (multi-rank) aggregation vs the same layer on the whole graph, forward and backward.

Two processes share cuda:0 under a gloo group — NCCL refuses two ranks on one device, so the halo rows
are staged through host memory by dist._a2a_rows; every kernel (pack, fused gather, GEMMs, halo-gradient
sum) is the product's.  The NCCL transport itself is exercised by ``bench.py --gpus N --check``."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
R, D, U = 5, 64, 250


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _base():
    from stargcn_b200 import synth
    return synth.make_bipartite(300, 200, 6000, n_levels=R, seed=3)


def _params():
    rs = np.random.RandomState(11)
    ws = [rs.uniform(-0.2, 0.2, (U, D)).astype(np.float32) for _ in range(R)]
    bs = [rs.uniform(-0.2, 0.2, (U,)).astype(np.float32) for _ in range(R)]
    return ws, bs


def _globals(world, base):
    rs = np.random.RandomState(5)
    x_item = rs.normal(size=(world * base["n_item"], D)).astype(np.float32)
    gout = rs.normal(size=(world * base["n_user"], U)).astype(np.float32)
    return x_item, gout


def _make_agg(ws, bs):
    from stargcn_b200.layers import MultiLinkGCNAggregator
    agg = MultiLinkGCNAggregator(units=U, num_links=R, act="leaky", ordinal_sharing=False, accum="sum", in_units=D).cuda()
    with torch.no_grad():
        for i in range(R):
            getattr(agg, f"weight{i}").copy_(torch.from_numpy(ws[i]))
            getattr(agg, f"bias{i}").copy_(torch.from_numpy(bs[i]))
    return agg


def _worker(rank, world, port, q, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from stargcn_b200 import dist as sgd, synth
        from stargcn_b200.graph import MultiLinkCSR
        base = _base()
        part = sgd.partitioned_layer_inputs(base, rank, world)
        indptr, cols, vals, sup = part["user"]
        plan = sgd.HaloPlan(cols, part["item_ranges"], rank, world, index_device="cpu", mode=mode)
        assert plan.mode == mode
        ep_l, ptr_l, sup_l, _ = synth.split_by_level(indptr, plan.local_cols, vals, sup, base["levels"])
        csr = MultiLinkCSR(ep_l, ptr_l, sup_l, n_nb=plan.n_ext, device="cuda")
        x_item, gout = _globals(world, base)
        lo, hi = part["item_ranges"][rank], part["item_ranges"][rank + 1]
        ulo, uhi = part["user_ranges"][rank], part["user_ranges"][rank + 1]
        x_local = torch.from_numpy(x_item[lo:hi]).cuda().requires_grad_(True)
        agg = _make_agg(*_params())
        out = agg(sgd.halo_exchange(x_local, plan), csr)
        out.backward(torch.from_numpy(gout[ulo:uhi]).cuda())
        sgd.allreduce_grads(list(agg.parameters()))
        q.put((rank, dict(out=out.detach().cpu().numpy(), gx=x_local.grad.cpu().numpy(),
                          gw=agg.weight2.grad.cpu().numpy(), gb=agg.bias2.grad.cpu().numpy(), n_halo=plan.n_halo)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("mode", ["alltoall", "allgather"])
def test_partitioned_layer_matches_whole_graph(mode):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=400) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    # the whole graph on one device
    from stargcn_b200 import dist as sgd, synth
    from stargcn_b200.graph import MultiLinkCSR
    base = _base()
    parts = [sgd.partitioned_layer_inputs(base, r, world) for r in range(world)]
    indptr = np.concatenate([[0]] + [np.diff(p["user"][0]) for p in parts]).cumsum().astype(np.int32)
    cols = np.concatenate([p["user"][1] for p in parts]).astype(np.int32)
    vals = np.concatenate([p["user"][2] for p in parts])
    sup = np.concatenate([p["user"][3] for p in parts])
    ep_l, ptr_l, sup_l, _ = synth.split_by_level(indptr, cols, vals, sup, base["levels"])
    x_item, gout = _globals(world, base)
    csr = MultiLinkCSR(ep_l, ptr_l, sup_l, n_nb=x_item.shape[0], device="cuda")
    agg = _make_agg(*_params())
    xg = torch.from_numpy(x_item).cuda().requires_grad_(True)
    out = agg(xg, csr)
    out.backward(torch.from_numpy(gout).cuda())
    out_h, gx_h = out.detach().cpu().numpy(), xg.grad.cpu().numpy()
    nu, ni = base["n_user"], base["n_item"]
    for r in range(world):
        assert got[r]["n_halo"] > 0
        assert rel_err(got[r]["out"], out_h[r * nu:(r + 1) * nu]) <= TOL
        assert rel_err(got[r]["gx"], gx_h[r * ni:(r + 1) * ni]) <= TOL      # includes gradients returned by peers
        assert rel_err(got[r]["gw"], agg.weight2.grad.cpu().numpy()) <= TOL   # all-reduced == whole-graph gradient
        assert rel_err(got[r]["gb"], agg.bias2.grad.cpu().numpy()) <= TOL


@pytest.mark.timeout(900)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="the NCCL transport needs two GPUs (run by `gpurun --gpus 2`)")
def test_partitioned_step_over_nccl_matches_whole_graph():
    """`bench.py --check` under torchrun: the partitioned step over NCCL (all-to-all and all-gather / reduce-scatter
    exchange on the weak construction, all-to-all on the nnz-balanced strong partition) and over this library's own
    NVLink peer-memory kernels (csrc/peer.cu; equal and nnz-balanced ranges, buffers reused across a forward-only pass
    and two training steps) against the whole graph."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(root, "bench.py"), "--gpus", str(n), "--check"],
                       capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"] and set(line["cases"]) == {"weak/alltoall", "weak/allgather", "strong/alltoall", "weak/peer", "strong/peer",
                                                      "weak/peer_sparse", "strong/peer_sparse"}
    for case in line["cases"].values():
        assert case["ok"] and max(case["out"], case["gx"], case["gw"], case["gb"]) <= 1e-5
