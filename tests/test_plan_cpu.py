"""Host-side computing plan (StackedHeterGCNLayers.gen_plan, mirror of mxgraph/layers/layers.py:260-337) and
the id-merging helpers it uses (mxgraph/graph.py:63-219 / graph_sampler.h:510-534 semantics): structure and
invariants on a small two-type graph with a stand-in for the reference's HeterGraph / CSRMat."""
import numpy as np

from stargcn_b200.hetergraph import merge_nodes, unordered_unique


def test_unordered_unique_first_appearance_order():
    ids = np.array([7, 3, 7, 9, 3, 1, 9, 7], np.int32)
    uniq, inv = unordered_unique(ids, return_inverse=True)
    assert uniq.tolist() == [7, 3, 9, 1]
    assert np.array_equal(uniq[inv], ids)
    assert unordered_unique(ids).tolist() == [7, 3, 9, 1]
    u, (a, b) = merge_nodes([ids[:3], ids[3:]])
    assert np.array_equal(u, uniq) and np.array_equal(np.concatenate([a, b]), inv)
    e, ei = unordered_unique(np.zeros(0, np.int32), return_inverse=True)
    assert e.size == 0 and ei.size == 0


from hostgraph import HostCSR as _HostCSR, HostGraph as _HostGraph  # noqa: E402


def _graph(seed=0, n_user=12, n_item=9, nnz=40, R=3):
    rs = np.random.RandomState(seed)
    flat = np.sort(rs.choice(n_user * n_item, nnz, replace=False))
    u, i = flat // n_item, flat % n_item
    levels = np.arange(1, R + 1).astype(np.float32)
    vals = levels[rs.randint(0, R, nnz)]
    uid, iid = np.arange(100, 100 + n_user, dtype=np.int32), np.arange(500, 500 + n_item, dtype=np.int32)

    def csr(r, c, n_r):
        order = np.lexsort((c, r))
        ptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=n_r))]).astype(np.int32)
        return ptr, c[order].astype(np.int32), vals[order]
    pu, cu, vu = csr(u, i, n_user)
    pi, ci, vi = csr(i, u, n_item)
    return _HostGraph({("user", "item"): _HostCSR(pu, cu, vu, levels, uid, iid),
                       ("item", "user"): _HostCSR(pi, ci, vi, levels, iid, uid)}), uid, iid, R


def test_gen_plan_structure_and_invariants():
    from stargcn_b200.layers import HeterGCNLayer, StackedHeterGCNLayers
    graph, uid, iid, R = _graph()
    mls = {("user", "item"): R, ("item", "user"): R}
    enc = StackedHeterGCNLayers()
    for _ in range(2):
        enc.add(HeterGCNLayer(meta_graph=graph.meta_graph, multi_link_structure=mls, agg_units=12, out_units=8,
                              agg_accum="sum", agg_act="leaky", out_act="leaky"))
    sel = {"user": np.array([103, 101, 103, 110], np.int32), "item": np.array([505, 505, 500], np.int32)}
    fanout = {("user", "item"): -1, ("item", "user"): -1}
    req_ids, plan = enc.gen_plan(graph, sel, graph_sampler_args=fanout, symm=True)
    assert len(plan) == 2 and req_ids is plan[0][0]
    # outermost depth: duplicates removed in order of first appearance, restore index brings them back
    top_ids, top_args = plan[1]
    for key in sel:
        row_inds, restore, entries = top_args[key]
        uniq = unordered_unique(sel[key])
        assert np.array_equal(top_ids[key][row_inds], uniq)             # local rows address the selected nodes
        assert np.array_equal(uniq[restore], sel[key])
    # every depth: end points are LOCAL indices into that depth's merged id list and map back to real neighbours
    for depth in (1, 0):
        ids, args = plan[depth]
        for src, (row_inds, restore, entries) in args.items():
            src_ids = ids[src][row_inds]
            for dst, (end_points, values, ind_ptr, support) in entries.items():
                mat = graph[src, dst]
                assert isinstance(end_points, list) and len(end_points) == R == len(ind_ptr)
                for lvl in range(R):
                    assert ind_ptr[lvl].shape[0] == len(src_ids) + 1 and ind_ptr[lvl][-1] == len(end_points[lvl])
                    for k, node in enumerate(src_ids):
                        got = ids[dst][end_points[lvl][ind_ptr[lvl][k]:ind_ptr[lvl][k + 1]]]
                        r = mat._row_of[int(node)]
                        sl = slice(mat.indptr[r], mat.indptr[r + 1])
                        want = mat.col_ids[mat.cols[sl]][mat.vals[sl] == mat.levels[lvl]]
                        assert np.array_equal(got, want)
        # the ids of depth d are exactly the nodes depth d-1 has to produce
        if depth == 1:
            for key in ids:
                assert np.array_equal(plan[0][0][key][plan[0][1][key][0]], ids[key])
