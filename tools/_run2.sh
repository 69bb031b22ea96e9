set -u
O=gpurun_out; mkdir -p $O
for v in 0 7 9 8 10; do
  echo "== bench SG_GATHER_SHAPE=$v"
  SG_GATHER_SHAPE=$v timeout 150 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline > "$O/b2_shape$v.json" 2> "$O/b2_shape$v.err"
  python tools/bench_summary.py < "$O/b2_shape$v.json"
done
echo "== parity subset with SHAPE=8 (pipelined kernel on every D=64 launch)"
SG_GATHER_SHAPE=8 timeout 300 python -m pytest tests/test_segops_gpu.py tests/test_layers_gpu.py tests/test_fullsize_gpu.py tests/test_stargcn_e2e_gpu.py -m gpu -x -q 2>&1 | tail -3
