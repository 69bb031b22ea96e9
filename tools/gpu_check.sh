#!/bin/bash
# One gpurun call: GPU parity suite, bench (default + A/B knobs), ncu launch list and one --set full capture.
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

el "bench default"
timeout 300 python bench.py --steps 100 --warmup 10 > $O/bench_default.json 2> $O/bench_default.err
echo "bench exit $?"; python tools/bench_summary.py < $O/bench_default.json

if [ "${1:-}" != "quick" ]; then
  for v in "SG_GATHER_SHAPE=6" "SG_PACK_FUSED=0" "SG_GATHER_SHAPE=4"; do
    el "bench $v"
    env $v timeout 150 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline > "$O/bench_$v.json" 2> "$O/bench_$v.err"
    python tools/bench_summary.py < "$O/bench_$v.json"
  done
  el "ncu launch list (eager, 5 steps)"
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_eager.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  el "cooperative gather: parity subset"
  SG_GATHER_SHAPE=4 timeout 300 python -m pytest tests/test_segops_gpu.py tests/test_layers_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > $O/pytest_coop.log 2>&1
  tail -3 $O/pytest_coop.log
  el "ncu --set full, gather kernels of one step"
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:gather_rows -s 12 -c 4 -o $O/prof_gather \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
el "done"
