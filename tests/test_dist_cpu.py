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


@pytest.mark.parametrize("world", [2, 3, 8])
def test_peer_sparse_layout_row_arithmetic(world):
    """dist.peer_sparse_layout (the pointer arithmetic behind the sparse-halo peer transport) for every rank of a
    random `world`: the forward blocks tile every receiver's table exactly as HaloPlan lays its halo out
    ([own rows ; rows from rank 0, 1, ... by id]) and the backward blocks tile every owner's staging in send-list
    order — simulated with numpy in place of the peer stores."""
    from stargcn_b200 import dist as sgd
    rs = np.random.RandomState(world)
    D = 4
    n_local = [int(v) for v in rs.randint(3, 30, size=world)]
    # requested rows: want[p][q] = sorted distinct local row ids of q that p fetches (p != q), some empty
    want = [[np.sort(rs.choice(n_local[q], size=rs.randint(0, n_local[q] + 1), replace=False)) if (p != q and rs.rand() > 0.2)
             else np.zeros(0, np.int64) for q in range(world)] for p in range(world)]
    counts = [dict(n_local=n_local[p], recv=[len(want[p][q]) for q in range(world)], send=[len(want[q][p]) for q in range(world)])
              for p in range(world)]
    lays = [sgd.peer_sparse_layout(counts, r) for r in range(world)]
    x = [rs.normal(size=(n_local[p], D)) for p in range(world)]
    # ---- forward: every rank packs the rows each peer asked for and stores them into that peer's table ----
    tables = [np.full((lays[0]["x_rows"], D), np.nan) for _ in range(world)]
    written = [np.zeros(lays[0]["x_rows"], np.int32) for _ in range(world)]
    for r in range(world):
        tables[r][:n_local[r]] = x[r]
        written[r][:n_local[r]] += 1
        send_cat = np.concatenate([want[p][r] for p in range(world)]).astype(np.int64)      # r's send list, by requesting rank
        assert lays[r]["n_send"] == send_cat.size and lays[r]["send_lo"][-1] == send_cat.size
        for p in range(world):
            a, b = lays[r]["send_lo"][p], lays[r]["send_lo"][p + 1]
            dst = lays[r]["x_dst"][p]
            tables[p][dst:dst + (b - a)] = x[r][send_cat[a:b]]
            written[p][dst:dst + (b - a)] += 1
    for p in range(world):
        n_ext = lays[p]["n_ext"]
        assert n_ext == n_local[p] + sum(counts[p]["recv"]) and n_ext <= lays[p]["x_rows"]
        assert np.all(written[p][:n_ext] == 1) and np.all(written[p][n_ext:] == 0)           # tiled exactly once
        expect = np.concatenate([x[p]] + [x[q][want[p][q]] for q in range(world)])
        assert np.array_equal(tables[p][:n_ext], expect)
    # ---- backward: every rank stores the gradient of each fetched slot into its owner's staging ----
    g_ext = [rs.normal(size=(lays[p]["n_ext"], D)) for p in range(world)]
    stage = [np.full((lays[0]["g_rows"], D), np.nan) for _ in range(world)]
    hits = [np.zeros(lays[0]["g_rows"], np.int32) for _ in range(world)]
    for r in range(world):
        hl = lays[r]["halo_lo"]
        assert hl[0] == 0 and hl[1] == n_local[r] and hl[-1] == lays[r]["n_ext"] and len(hl) == world + 2
        for q in range(world):
            a, b = hl[1 + q], hl[2 + q]
            dst = lays[r]["g_dst"][q]
            stage[q][dst:dst + (b - a)] = g_ext[r][a:b]
            hits[q][dst:dst + (b - a)] += 1
    for q in range(world):
        n_send = lays[q]["n_send"]
        assert np.all(hits[q][:n_send] == 1) and np.all(hits[q][n_send:] == 0)
        send_cat = np.concatenate([want[p][q] for p in range(world)]).astype(np.int64)
        got = g_ext[q][:n_local[q]].copy()
        np.add.at(got, send_cat, stage[q][:n_send])                                            # the owner's sorted-transpose sum
        expect = g_ext[q][:n_local[q]].copy()
        for p in range(world):
            off = n_local[p] + sum(counts[p]["recv"][:q])
            np.add.at(expect, want[p][q], g_ext[p][off:off + len(want[p][q])])
        assert np.allclose(got, expect, rtol=0, atol=1e-12)


def test_halo_plan_modes_without_a_process_group():
    """Mode / layout choice of HaloPlan where no collective is needed (exchange_fn given or forced modes)."""
    from stargcn_b200 import dist as sgd
    ranges = np.array([0, 10, 20, 30])
    cols_dense = np.arange(30)                       # rank 1 references every row of every rank
    cols_sparse = np.array([10, 11, 12, 0, 29])      # ... or 2 of the 20 remote rows
    echo = lambda req: [np.zeros(0, np.int32) for _ in req]
    for mode, cols, want in (("peer_dense", cols_sparse, ("peer", True)), ("peer_sparse", cols_dense, ("peer", False)),
                             ("alltoall", cols_dense, ("alltoall", False)), ("allgather", cols_sparse, ("allgather", True)),
                             ("auto", cols_dense, ("allgather", True)), ("auto", cols_sparse, ("alltoall", False)),
                             ("nccl", cols_dense, ("allgather", True)), ("peer", cols_dense, ("peer", True)),
                             ("peer", cols_sparse, ("peer", False))):
        plan = sgd.HaloPlan(cols, ranges, 1, 3, exchange_fn=echo, mode=mode)
        assert (plan.mode, plan.dense) == want, (mode, plan.mode, plan.dense)
        if plan.dense:
            assert plan.n_ext == 30 and np.array_equal(plan.local_cols, cols)
        else:
            assert plan.n_local == 10 and plan.n_ext == 10 + np.unique(cols[(cols < 10) | (cols >= 20)]).size
    with pytest.raises(ValueError):
        sgd.HaloPlan(cols_dense, ranges, 1, 3, exchange_fn=echo, mode="bogus")
    with pytest.raises(ValueError):                   # unequal blocks cannot use the NCCL all-gather
        sgd.HaloPlan(np.arange(25), np.array([0, 10, 25]), 0, 2, exchange_fn=echo, mode="allgather")
    with pytest.raises(ValueError):                   # one NVSwitch box: at most 8 ranks over peer memory
        sgd.HaloPlan(np.arange(9), np.arange(10), 0, 9, exchange_fn=echo, mode="peer_dense")
    with pytest.raises(ValueError):
        sgd.HaloPlan(np.array([31]), ranges, 1, 3, exchange_fn=echo, mode="alltoall")
