"""Developer tool: torch.profiler table of one bench step (run under torchrun for N>1).
    torchrun --nproc-per-node 2 tools/prof_step.py --gpus 2"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    args = bench.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # reuse bench's setup by monkeypatching the timed loop: run_gpu_arm builds everything; we only need step()
    captured = {}
    orig_sampler = bench.ClockSampler.start

    def grab(self):
        import inspect
        fr = inspect.currentframe().f_back
        captured["step"] = fr.f_locals["step"]
        orig_sampler(self)
    bench.ClockSampler.start = grab
    args.steps, args.warmup, args.no_e2e = 3, 3, True
    bench.run_gpu_arm(args, rank, world, local)
    step = captured["step"]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    t_cpu = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 10
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        print(f"cpu issue time per step {t_cpu * 1e3:.3f} ms, wall per step {t_all * 1e3:.3f} ms")
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
        print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=25, max_name_column_width=60))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
