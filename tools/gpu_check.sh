#!/bin/bash
# One gpurun call: GPU parity suite, default bench, ncu launch list and one --set full capture of the gather kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1

el "pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest exit $?" | tee -a $O/pytest.log
tail -4 $O/pytest.log

el "smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

el "bench default"
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench exit $?"; python tools/bench_summary.py < $O/bench_default.json

if [ "${1:-}" != "quick" ]; then
  el "ncu launch list (eager, 5 steps)"
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_eager.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  el "ncu --set full, gather kernels of one step"
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:gather_rows -s 12 -c 4 -o $O/prof_gather \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
el "done"
