#!/bin/bash
# One gpurun call: GPU parity suite, smoke, default bench, A/B benches.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [quick|ncu]
set -u
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1

el "pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --tb=short > $O/pytest.log 2>&1
echo "pytest exit $?" | tee -a $O/pytest.log
grep -vE "Warning|warn|^$|^  " $O/pytest.log | tail -15

el "smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

el "bench default"
timeout 500 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench exit $?"; python tools/bench_summary.py < $O/bench_default.json; tail -3 $O/bench_default.err

if [ "${1:-}" != "quick" ]; then
  el "A/B: GEMM operands split in the kernel"
  timeout 300 python bench.py --inkernel-split --no-e2e --no-cpu-baseline > $O/bench_inkernel.json 2> $O/bench_inkernel.err
  python tools/bench_summary.py < $O/bench_inkernel.json
  el "gather sweep: default vs staged (TMA bulk copy) variant"
  timeout 300 python tools/sweep_gather.py "gather_variant=0" "gather_variant=1" > $O/sweep_gather.log 2>&1; tail -6 $O/sweep_gather.log
fi
if [ "${1:-}" == "ncu" ]; then
  el "ncu launch list (eager, 5 steps)"
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_eager.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  el "ncu --set full, gather + gemm kernels of one step"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gather_rows|tf32x3' -s 20 -c 10 -o $O/prof_step \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
el "done"
