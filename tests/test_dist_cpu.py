"""Host-side logic of the node-partitioned path (star-gcn_b200/dist.py), world_size 2 and 3 over gloo on
CPU tensors: partition ranges, halo index plans (who fetches which rows), column rewriting, and the
consistency of the synthetic partitioned workload.  The row traffic itself (CUDA kernels + NCCL) is
covered by tests/test_dist_gpu.py and bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _small_base(seed=0):
    from stargcn_b200 import synth
    return synth.make_bipartite(60, 40, 700, n_levels=5, seed=seed)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stargcn_b200 import dist as sgd
        base = _small_base()
        part = sgd.partitioned_layer_inputs(base, rank, world)
        rs = np.random.RandomState(1)
        x_item_global = rs.normal(size=(world * base["n_item"], 8)).astype(np.float32)   # same on every rank
        x_user_global = rs.normal(size=(world * base["n_user"], 8)).astype(np.float32)
        res = {}
        for side, ranges, xg in (("user", part["item_ranges"], x_item_global), ("item", part["user_ranges"], x_user_global)):
            indptr, cols, vals, sup = part[side]
            plan = sgd.HaloPlan(cols, ranges, rank, world, index_device="cpu", mode="alltoall")
            # dense halo -> the collectively chosen mode is the all-gather one: global ids index the gathered table
            auto = sgd.HaloPlan(cols, ranges, rank, world, index_device="cpu", mode="auto")
            ok_auto = auto.mode == "allgather" and auto.n_ext == xg.shape[0] and np.array_equal(auto.local_cols, cols)
            # the peer-memory mode keeps global ids as well (any contiguous ranges; the kernels need CUDA, the plan does not)
            peer = sgd.HaloPlan(cols, ranges, rank, world, index_device="cpu", mode="peer")
            ok_auto = (ok_auto and peer.mode == "peer" and peer.dense and peer.n_ext == xg.shape[0]
                       and np.array_equal(peer.local_cols, cols) and peer.n_local == ranges[rank + 1] - ranges[rank])
            # ... and its sparse layout is the all-to-all plan: [own rows ; deduplicated halo rows by owner, by id]
            sparse = sgd.HaloPlan(cols, ranges, rank, world, index_device="cpu", mode="peer_sparse")
            ok_auto = (ok_auto and sparse.mode == "peer" and not sparse.dense and np.array_equal(sparse.local_cols, plan.local_cols)
                       and sparse.send_counts == plan.send_counts and sparse.recv_counts == plan.recv_counts)
            lo, hi = ranges[rank], ranges[rank + 1]
            x_local = torch.from_numpy(xg[lo:hi])
            # row exchange emulated with a gloo all-to-all on CPU tensors (test stand-in for pack kernel + NCCL)
            send_cat = np.concatenate(plan.send_idx) if sum(plan.send_counts) else np.zeros(0, np.int64)
            send = x_local[torch.from_numpy(send_cat.astype(np.int64))]
            halo = torch.empty((plan.n_halo, 8))
            dist.all_to_all_single(halo, send, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts)
            x_ext = torch.cat([x_local, halo]).numpy()
            ok_rows = np.array_equal(x_ext[plan.local_cols], xg[cols])          # rewritten ids address the right rows
            ok_sorted = all(np.all(np.diff(r) > 0) for r in plan.recv_ids if r.size > 1)
            res[side] = dict(ok_rows=bool(ok_rows and ok_auto), ok_sorted=bool(ok_sorted), n_halo=plan.n_halo, n_local=plan.n_local,
                             send_counts=plan.send_counts, recv_counts=plan.recv_counts, nnz=int(cols.size),
                             edges=set(zip(np.repeat(np.arange(len(indptr) - 1) + (part["user_ranges"] if side == "user" else part["item_ranges"])[rank],
                                                     np.diff(indptr)).tolist(), cols.tolist())))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_plan_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        for side in ("user", "item"):
            assert out[r][side]["ok_rows"] and out[r][side]["ok_sorted"]
            # what r sends to q is what q receives from r
            for qq in range(world):
                assert out[r][side]["send_counts"][qq] == out[qq][side]["recv_counts"][r]
            assert out[r][side]["send_counts"][r] == 0
    # the two directions describe the same global edge set
    user_edges = set().union(*[out[r]["user"]["edges"] for r in range(world)])
    item_edges = set().union(*[{(u, i) for (i, u) in out[r]["item"]["edges"]} for r in range(world)])
    assert user_edges == item_edges
    base = _small_base()
    assert len(user_edges) == world * base["nnz"]


def test_ranges():
    from stargcn_b200 import dist as sgd
    r = sgd.contiguous_ranges(10, 3)
    assert r[0] == 0 and r[-1] == 10 and np.all(np.diff(r) >= 3)
    deg = np.array([100, 1, 1, 1, 1, 100, 1, 1])
    b = sgd.balanced_ranges(deg, 2)
    assert b[0] == 0 and b[-1] == 8
    left = deg[:b[1]].sum()
    assert abs(left - deg.sum() / 2) <= 100


def test_halo_plan_single_rank_is_identity():
    from stargcn_b200 import dist as sgd
    cols = np.array([3, 1, 1, 0, 2], np.int64)
    plan = sgd.HaloPlan(cols, np.array([0, 4]), 0, 1)
    assert plan.n_halo == 0 and plan.n_local == 4
    assert np.array_equal(plan.local_cols, cols.astype(np.int32))
    with pytest.raises(ValueError):
        sgd.HaloPlan(np.array([5]), np.array([0, 4]), 0, 1)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_strong_partition_covers_the_graph_once(world):
    """bench.partition_sides(scaling='strong'): the nnz-balanced node ranges tile both sides, the per-rank CSR slices
    put together are the whole graph, and the per-rank edge counts are balanced ('balanced by nnz')."""
    import bench
    from stargcn_b200 import synth
    base = synth.make_bipartite(400, 150, 9000, n_levels=5, seed=2)
    for side, key in (("user", "u2i"), ("item", "i2u")):
        rows, edges, los = 0, [], []
        for rank in range(world):
            ps = bench.partition_sides(base, rank, world, "strong")[side]
            indptr, cols, vals, sup = ps["csr"]
            assert indptr[0] == 0 and indptr[-1] == cols.size == vals.size == sup.size
            assert ps["n_dst"] == len(indptr) - 1
            lo = ps["dst_lo"]
            p0 = base[key]["indptr"][lo]
            assert np.array_equal(cols, base[key]["cols"][p0:p0 + cols.size])
            assert np.array_equal(sup, base[key]["support"][p0:p0 + cols.size])
            assert ps["nb_ranges"][0] == 0 and ps["nb_ranges"][-1] == (base["n_item"] if side == "user" else base["n_user"])
            rows += ps["n_dst"]; edges.append(int(cols.size)); los.append(lo)
        assert rows == (base["n_user"] if side == "user" else base["n_item"])
        assert sum(edges) == base["nnz"] and los == sorted(los)
        max_deg = int(np.diff(base[key]["indptr"]).max())
        assert max(edges) - min(edges) <= 2 * max_deg + 1      # balanced up to one row's worth of edges per cut
