"""Developer tool: time the pieces of the per-step plan (re)build on the ML-10M shape: schedules, transposed operands by
radix sort vs read off the reverse plan, load_lists_ from pinned memory."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import stargcn_b200  # noqa: F401,E402
from stargcn_b200.graph import MultiLinkCSR  # noqa: E402


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


wl = bench.load_workload(os.environ.get("SWEEP_WORKLOAD", "ml-10m"))
dev = torch.device("cuda", 0)
u = MultiLinkCSR(*wl["user"], n_nb=wl["n_item"], device=dev).prepare()
i = MultiLinkCSR(*wl["item"], n_nb=wl["n_user"], device=dev).prepare()
print(f"sort path   : user rebuild_ {t(u.rebuild_):.3f} ms, item rebuild_ {t(i.rebuild_):.3f} ms")
print(f"  schedules only: user {t(lambda: (u._sched.rebuild_(), u._t_sched.rebuild_())):.3f} ms, item {t(lambda: (i._sched.rebuild_(), i._t_sched.rebuild_())):.3f} ms")
print(f"  transposed only (sort): user {t(lambda: u._build_transposed(u._t)):.3f} ms, item {t(lambda: i._build_transposed(i._t)):.3f} ms")
u2 = u
pin = lambda lst: [torch.from_numpy(a).pin_memory() for a in lst]
lists = [pin(x) for x in wl["user"][:3]]
print(f"load_lists_ (user side, pinned -> device, {u2.h2d_bytes / 1e6:.0f} MB): {t(lambda: u2.load_lists_(*lists)):.3f} ms")
