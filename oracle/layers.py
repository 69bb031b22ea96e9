"""TEST INFRASTRUCTURE ONLY — numpy restatement of the layer-level maths of the hot path.

  multilink_aggregator_*   MultiLinkGCNAggregator.hybrid_forward in the REFERENCE's operator
                           order — per level FullyConnected then seg_weighted_pool, then
                           concat | add_n, then activation (mxgraph/layers/aggregators.py:111-163)
                           — with the pooling done by the pinned C oracle (oracle/segops.py)
  heter_layer_forward      HeterGCNLayer.forward_single for one neighbour type
                           (mxgraph/layers/layers.py:147-187): aggregator -> Dense -> act
  get_embed / decoder / recon_loss / rating head
                           experiments/STAR-GCN.py:264-300, 226-246 + 441-454, 618-628, 428-438

Parity pin status: the seg_weighted_pool part is pinned (oracle/_ref); FullyConnected,
LeakyReLU, Embedding/take and the losses live in un-vendored Apache MXNet 1.5.x, whose
documented semantics are restated here (y = x W^T + b with W (units, in); LeakyReLU(0.1):
x > 0 ? x : 0.1 x; take mode clip; L2Loss = 0.5 (p - y)^2) — "parity unpinned" for those.
Dropout is the identity here (parity runs use rate 0 / eval mode).
"""
import numpy as np

from . import segops

LEAKY = 0.1


def act_fwd(x, act):
    if act in (None, "identity"):
        return x
    if act == "leaky":
        return np.where(x > 0, x, np.asarray(LEAKY, x.dtype) * x)
    if act == "relu":
        return np.maximum(x, 0)
    raise NotImplementedError(act)


def act_bwd(pre, g, act):
    if act in (None, "identity"):
        return g
    if act == "leaky":
        return np.where(pre > 0, g, np.asarray(LEAKY, g.dtype) * g)
    if act == "relu":
        return np.where(pre > 0, g, 0)
    raise NotImplementedError(act)


def _effective_params(weights, biases, ordinal_sharing):
    if not ordinal_sharing:
        return list(weights), list(biases)
    ws, bs = [weights[0]], [biases[0]]
    for w, b in zip(weights[1:], biases[1:]):
        ws.append(ws[-1] + w)
        bs.append(bs[-1] + b)
    return ws, bs


def _pool64(h, s, e, ptr):
    n_seg = len(ptr) - 1
    out = np.zeros((n_seg, h.shape[1]), np.float64)
    seg = np.repeat(np.arange(n_seg), np.diff(ptr))
    np.add.at(out, seg, s[:, None].astype(np.float64) * h[e].astype(np.float64))
    return out


def multilink_aggregator_forward(x, weights, biases, end_points_l, indptr_l, support_l, accum="sum", act="leaky",
                                 ordinal_sharing=False, fp64=False):
    """Returns (out, pre_activation).  fp64=True gives the error-budget reference."""
    dt = np.float64 if fp64 else np.float32
    x = np.asarray(x, dt)
    ws, bs = _effective_params([np.asarray(w, dt) for w in weights], [np.asarray(b, dt) for b in biases],
                               ordinal_sharing)
    outs = []
    for w, b, e, ptr, s in zip(ws, bs, end_points_l, indptr_l, support_l):
        nnz = int(ptr[-1])
        e, s = np.asarray(e[:nnz], np.int32), np.asarray(s[:nnz], np.float32)
        h = x @ w.T + b                                                        # FullyConnected, aggregators.py:141
        if fp64:
            outs.append(_pool64(h, s, e, ptr))
        else:
            outs.append(segops.seg_weighted_pool(h[None], s[None], e, ptr)[0])  # aggregators.py:146
    if len(outs) == 1:
        pre = outs[0]
    elif accum == "stack":
        pre = np.concatenate(outs, axis=1)
    else:
        pre = outs[0].copy()
        for o in outs[1:]:
            pre = pre + o                                                       # add_n
    return act_fwd(pre, act), pre


def multilink_aggregator_backward(x, weights, biases, end_points_l, indptr_l, support_l, gout, pre, accum="sum",
                                  act="leaky", ordinal_sharing=False, fp64=False):
    """Gradients (gx, [gW_r], [gb_r]) of sum(out * gout); chain rule through the reference graph with
    the data-gradient scatter of seg_op.cc:209-240."""
    dt = np.float64 if fp64 else np.float32
    x = np.asarray(x, dt)
    ws, _ = _effective_params([np.asarray(w, dt) for w in weights], [np.asarray(b, dt) for b in biases],
                              ordinal_sharing)
    gz = act_bwd(pre, np.asarray(gout, dt), act)
    R = len(ws)
    gx = np.zeros_like(x)
    gws, gbs = [], []
    for r, (w, e, ptr, s) in enumerate(zip(ws, end_points_l, indptr_l, support_l)):
        nnz = int(ptr[-1])
        e, s = np.asarray(e[:nnz], np.int32), np.asarray(s[:nnz], np.float32)
        g_r = gz if (accum == "sum" or R == 1) else gz[:, r * w.shape[0]:(r + 1) * w.shape[0]]
        g_r = np.ascontiguousarray(g_r)
        if fp64:
            gh = np.zeros((x.shape[0], w.shape[0]), np.float64)
            seg = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
            np.add.at(gh, e, s[:, None].astype(np.float64) * g_r[seg])
        else:
            gh = segops.seg_weighted_pool_bwd_data(g_r[None], s[None], e, ptr, x.shape[0])[0]
        gws.append(gh.T @ x)
        gbs.append(gh.sum(axis=0))
        gx = gx + gh @ w
    if ordinal_sharing:  # W_r = sum_{i<=r} weight_i  ->  grad(weight_i) = sum_{r>=i} gW_r
        for i in range(R - 2, -1, -1):
            gws[i] = gws[i] + gws[i + 1]
            gbs[i] = gbs[i] + gbs[i + 1]
    return gx, gws, gbs


def dense(x, w, b=None):
    y = x @ w.T
    return y if b is None else y + b


def heter_layer_forward(x_nb, agg_params, out_w, out_b, lists, accum="sum", act="leaky", out_act="leaky",
                        fp64=False):
    """One source node type with one neighbour type (bipartite): aggregator -> Dense(out) -> act."""
    weights, biases = agg_params
    h, _ = multilink_aggregator_forward(x_nb, weights, biases, *lists, accum=accum, act=act, fp64=fp64)
    dt = np.float64 if fp64 else np.float32
    return act_fwd(dense(h, np.asarray(out_w, dt), np.asarray(out_b, dt)), out_act)


# ---- decoder side (experiments/STAR-GCN.py) ----
def get_embed(table, node_ids, embed_noise=None, use_mask=True):
    """:264-300 — ids' = noise[ids]; mask = ids' != -1; E[ids' * mask] * mask."""
    ids = np.asarray(node_ids, np.int64)
    if use_mask:
        ids = np.asarray(embed_noise, np.int64)[ids]
        mask = ids != -1
        ids = ids * mask
    emb = table[ids]
    if use_mask:
        emb = emb * mask[:, None].astype(table.dtype)
    return emb


def get_embed_backward(table_shape, node_ids, gout, embed_noise=None, use_mask=True):
    ids = np.asarray(node_ids, np.int64)
    g = np.asarray(gout)
    if use_mask:
        ids = np.asarray(embed_noise, np.int64)[ids]
        mask = ids != -1
        ids = ids * mask
        g = g * mask[:, None].astype(g.dtype)
    gt = np.zeros(table_shape, g.dtype)
    np.add.at(gt, ids, g)
    return gt


def embed_map(h, idx, w0, b0, w1, b1, act="leaky"):
    """:441-454 — Dense -> act -> Dense applied to take(h, idx)."""
    z = dense(h[np.asarray(idx, np.int64)], w0, b0)
    return dense(act_fwd(z, act), w1, b1)


def recon_loss(gt_emb, pred_emb):
    """:625 — mean over nodes of the summed squared error."""
    d = gt_emb - pred_emb
    return (d * d).sum(axis=-1).mean()


def rating_head(h_user, h_item, idx_user, idx_item, wu, bu, wi, bi):
    """:428-438 + InnerProductLayer (layers.py:217-222) without mid map."""
    u = dense(h_user[np.asarray(idx_user, np.int64)], wu, bu)
    v = dense(h_item[np.asarray(idx_item, np.int64)], wi, bi)
    return (u * v).sum(axis=1, keepdims=True)


def l2_loss(pred, label):
    """gluon.loss.L2Loss: 0.5 * (pred - label)^2, mean over the batch (:611-616)."""
    return (0.5 * (pred.reshape(-1) - label.reshape(-1)) ** 2).mean()
