"""Developer tool: one full STAR-GCN training iteration (BASELINE.json configs[1] ML-100k, [2] ML-1M, [4] Douban-shaped:
two stacked blocks + masked-embedding reconstruction + rating head, D=64, U=250, O=75) end to end from PINNED HOST batch
arrays to the loss read-back, three ways:
   host plans    StarGCN.forward on a host graph object (numpy gen_plan + per-entry uploads), eager
   device plans  StarGCN.forward on a DeviceHeterGraph with the batch edges removed on the device, eager
   static graph  StaticGraphStep: batch edges masked on static whole-graph plans, ONE CUDA graph + fused clip/Adam
    python tools/bench_model.py [ml-100k|douban|ml-1m] [iterations]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import stargcn_b200  # noqa: F401,E402
from stargcn_b200 import devgraph, synth  # noqa: E402
from stargcn_b200.devgraph import DeviceCSRMat, DeviceHeterGraph  # noqa: E402
from stargcn_b200.model import StarGCN  # noqa: E402
from stargcn_b200.optim import FusedAdam  # noqa: E402
from stargcn_b200.static_step import StaticGraphStep  # noqa: E402


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "ml-100k"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    R, D = 5, 64
    n_user, n_item, n_edges, _, _ = synth.SHAPES[shape]
    B = {"ml-100k": 10_000, "douban": 10_000, "ml-1m": 100_000}[shape]      # TRAIN.RATING_BATCH_SIZE of the shipped cfgs
    g = synth.make_bipartite(n_user, n_item, n_edges + B, R, seed=1000)     # the train graph still holds the batch edges
    dg = DeviceHeterGraph.from_synth(g)
    rs = np.random.RandomState(0)
    n_rec = {"user": n_user // 10, "item": n_item // 10}
    batches = []
    for _ in range(8):
        pick = rs.choice(g["nnz"], B, replace=False)
        noise = {"user": np.arange(n_user, dtype=np.int32), "item": np.arange(n_item, dtype=np.int32)}
        recon = {}
        for k, n in (("user", n_user), ("item", n_item)):
            perm = rs.permutation(n)
            recon[k] = perm[:n_rec[k]].astype(np.int32)
            noise[k][perm[:n_rec[k] // 2]] = -1
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        batches.append(dict(pairs=pin(np.stack([g["u2i"]["rows"][pick], g["u2i"]["cols"][pick]]).astype(np.int32)),
                            ratings=pin(g["u2i"]["vals"][pick].astype(np.float32)),
                            noise={k: pin(v) for k, v in noise.items()}, recon={k: pin(v) for k, v in recon.items()}))
    torch.manual_seed(0)
    mls = {("user", "item"): R, ("item", "user"): R}
    model = StarGCN(dg.meta_graph, mls, {"user": n_user, "item": n_item}, "user", "item", embed_units=D, agg_units=250,
                    out_units=75, n_blocks=2, mid_map=64, agg_accum="sum", act="leaky").cuda()
    fan = {("user", "item"): -1, ("item", "user"): -1}
    mean, std, lam = 3.5, 1.1, 0.1

    def removed(b):
        ui, iu = dg["user", "item"], dg["item", "user"]
        pu, pi = b["pairs"][0].cuda(non_blocking=True), b["pairs"][1].cuda(non_blocking=True)
        ru, ci = ui.rows_of(pu), ui.cols_of(pi)
        return DeviceHeterGraph(dg.meta_graph, {("user", "item"): DeviceCSRMat(ui.csr.remove_edges(ru, ci), ui.row_ids, ui.col_ids),
                                                ("item", "user"): DeviceCSRMat(iu.csr.remove_edges(ci, ru), iu.row_ids, iu.col_ids)})

    def eager_device(b, opt):
        red = removed(b)
        model.zero_grad(set_to_none=True)
        pr, pe, gt = model(red, b["pairs"].numpy(), {k: v.numpy() for k, v in b["noise"].items()},
                           {k: v.numpy() for k, v in b["recon"].items()}, fan)
        loss = model.loss(pr, pe, gt, b["ratings"].cuda(non_blocking=True), mean, std, lam)
        loss.backward()
        if opt is not None:
            opt.clip_global_norm(1.0); opt.step()
        return float(loss.item())

    eager_device(batches[0], None)                         # materialise deferred shapes
    opt = FusedAdam(list(model.parameters()), learning_rate=2e-3)
    out = dict(shape=shape, users=n_user, items=n_item, train_edges=g["nnz"], batch=B, recon=n_rec)

    def timed(fn, n):
        for k in range(3):
            fn(batches[k % len(batches)])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(n):
            fn(batches[k % len(batches)])
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    out["device_plans_eager_ms"] = timed(lambda b: eager_device(b, opt), max(3, iters // 5))

    step = StaticGraphStep(model, dg, B, n_rec, rating_mean=mean, rating_std=std, recon_lambda=lam).capture()

    def static(b):
        loss = step(b["pairs"], b["ratings"], b["noise"], b["recon"])
        opt.clip_global_norm(1.0); opt.step()
        return float(loss.item())

    out["static_graph_ms"] = timed(static, iters)
    l0 = static(batches[0])
    for _ in range(20):
        for b in batches:
            last = static(b)
    out["loss_first_vs_after_160_steps"] = [l0, static(batches[0])]
    out["iterations_per_s_static"] = 1e3 / out["static_graph_ms"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
