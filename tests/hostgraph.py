"""TEST INFRASTRUCTURE — a host stand-in for the slice of mxgraph.graph.HeterGraph / CSRMat that
StackedHeterGCNLayers.gen_plan touches: ``graph.meta_graph`` and ``graph[src, dst].sample_neighbors``
(full neighbourhood; node ids = column ids of the stored matrix)."""
import numpy as np


class HostCSR:
    def __init__(self, indptr, cols, vals, levels, row_ids, col_ids, support=None):
        self.indptr, self.cols, self.vals, self.levels = indptr, cols, vals, levels
        self.row_ids, self.col_ids = row_ids, col_ids
        self.support = np.full(cols.size, 0.5, np.float32) if support is None else support
        self._row_of = {int(r): k for k, r in enumerate(row_ids)}

    def sample_neighbors(self, src_ids=None, symm=True, use_multi_link=True, num_neighbors=None):
        rows = np.array([self._row_of[int(i)] for i in src_ids], np.int64)
        lens = (self.indptr[rows + 1] - self.indptr[rows]).astype(np.int64)
        ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        pos = (np.repeat(self.indptr[rows].astype(np.int64) - ptr[:-1], lens) + np.arange(ptr[-1])).astype(np.int64)
        ep_ids, vals, sup = self.col_ids[self.cols[pos]], self.vals[pos], self.support[pos]
        if not use_multi_link:
            return ep_ids, vals, ptr, sup
        seg = np.repeat(np.arange(len(rows)), lens)
        ep_l, val_l, ptr_l, sup_l = [], [], [], []
        for lv in self.levels:
            m = vals == lv
            ep_l.append(ep_ids[m]); val_l.append(vals[m]); sup_l.append(sup[m])
            ptr_l.append(np.concatenate([[0], np.cumsum(np.bincount(seg[m], minlength=len(rows)))]).astype(np.int32))
        return ep_l, val_l, ptr_l, sup_l


class HostGraph:
    def __init__(self, mats, user="user", item="item"):
        self._mats = mats
        self.meta_graph = {user: {item: "rating"}, item: {user: "rev_rating"}}

    def __getitem__(self, key):
        return self._mats[key]


def from_synth(g, user="user", item="item"):
    """HostGraph over a stargcn_b200.synth.make_bipartite graph (node ids = 0..N-1 on each side)."""
    uid, iid = np.arange(g["n_user"], dtype=np.int32), np.arange(g["n_item"], dtype=np.int32)
    u2i, i2u = g["u2i"], g["i2u"]
    return HostGraph({(user, item): HostCSR(u2i["indptr"], u2i["cols"], u2i["vals"], g["levels"], uid, iid, u2i["support"]),
                      (item, user): HostCSR(i2u["indptr"], i2u["cols"], i2u["vals"], g["levels"], iid, uid, i2u["support"])},
                     user, item)
