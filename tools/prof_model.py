"""Developer tool: kernel table of ONE static-graph STAR-GCN training iteration (tools/bench_model.py's workload).
    python tools/prof_model.py [ml-100k|douban|ml-1m]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import stargcn_b200  # noqa: F401,E402
from stargcn_b200 import synth  # noqa: E402
from stargcn_b200.devgraph import DeviceHeterGraph  # noqa: E402
from stargcn_b200.model import StarGCN  # noqa: E402
from stargcn_b200.optim import FusedAdam  # noqa: E402
from stargcn_b200.static_step import StaticGraphStep  # noqa: E402


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "ml-100k"
    R, D = 5, 64
    n_user, n_item, n_edges, _, _ = synth.SHAPES[shape]
    B = {"ml-100k": 10_000, "douban": 10_000, "ml-1m": 100_000}[shape]
    g = synth.make_bipartite(n_user, n_item, n_edges + B, R, seed=1000)
    dg = DeviceHeterGraph.from_synth(g)
    rs = np.random.RandomState(0)
    n_rec = {"user": n_user // 10, "item": n_item // 10}
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    pick = rs.choice(g["nnz"], B, replace=False)
    noise = {"user": np.arange(n_user, dtype=np.int32), "item": np.arange(n_item, dtype=np.int32)}
    recon = {}
    for k, n in (("user", n_user), ("item", n_item)):
        perm = rs.permutation(n)
        recon[k] = perm[:n_rec[k]].astype(np.int32)
        noise[k][perm[:n_rec[k] // 2]] = -1
    b = dict(pairs=pin(np.stack([g["u2i"]["rows"][pick], g["u2i"]["cols"][pick]]).astype(np.int32)),
             ratings=pin(g["u2i"]["vals"][pick].astype(np.float32)),
             noise={k: pin(v) for k, v in noise.items()}, recon={k: pin(v) for k, v in recon.items()})
    torch.manual_seed(0)
    mls = {("user", "item"): R, ("item", "user"): R}
    model = StarGCN(dg.meta_graph, mls, {"user": n_user, "item": n_item}, "user", "item", embed_units=D, agg_units=250,
                    out_units=75, n_blocks=2, mid_map=64, agg_accum="sum", act="leaky").cuda()
    step = StaticGraphStep(model, dg, B, n_rec, rating_mean=3.5, rating_std=1.1, recon_lambda=0.1)
    step(b["pairs"], b["ratings"], b["noise"], b["recon"], eager=True)       # materialise
    opt = FusedAdam(list(model.parameters()), learning_rate=2e-3)
    step.capture()

    def it():
        loss = step(b["pairs"], b["ratings"], b["noise"], b["recon"])
        opt.clip_global_norm(1.0); opt.step()
        return float(loss.item())

    for _ in range(5):
        it()
    torch.cuda.synchronize()
    # where the wall time goes: host-side load, graph launch, optimiser, read-back
    marks = []
    for _ in range(20):
        t0 = time.perf_counter(); step.load(b["pairs"], b["ratings"], b["noise"], b["recon"])
        t1 = time.perf_counter(); step._graph()
        t2 = time.perf_counter(); opt.clip_global_norm(1.0); opt.step()
        t3 = time.perf_counter(); float(step.loss.item())
        t4 = time.perf_counter()
        marks.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    m = np.array(marks).mean(0) * 1e3
    print(f"host ms: load {m[0]:.3f}  graph launch {m[1]:.3f}  optimiser issue {m[2]:.3f}  wait for loss {m[3]:.3f}  total {m.sum():.3f}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        step._graph()
    e1.record(); torch.cuda.synchronize()
    print(f"graph replay alone: {e0.elapsed_time(e1) / 20:.3f} ms")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for _ in range(5):
            it()
        torch.cuda.synchronize()
    ka = [e for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(e.device_time_total for e in ka)
    n = sum(e.count for e in ka)
    print(f"device kernels+copies per iteration: {n / 5:.0f}, summed device time {tot / 5 / 1e3:.3f} ms")
    for e in sorted(ka, key=lambda e: -e.device_time_total)[:28]:
        print(f"  {e.key[:86]:86s} {e.count / 5:6.1f} x {e.device_time_total / e.count:7.1f} us = {e.device_time_total / 5:8.1f} us")


if __name__ == "__main__":
    main()
