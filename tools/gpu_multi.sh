#!/bin/bash
# One `gpurun --gpus N` call: parity check of every exchange transport + the partitioned bench.
# Usage: bash tools/gpu_multi.sh N [peertest|check|weak|nccl|a2a|psparse|strong|strongnccl|sm|ce|pushK ...]
#   weak / strong = default transport (peer memory); nccl / strongnccl / a2a = the NCCL collectives; psparse = peer memory
#   with the sparse-halo layout forced; sm / ce = all-gather by the store kernel / copy engines; pushK = store-kernel grid
set -u
N=${1:-2}; shift || true
WHAT=${*:-check weak a2a strong}
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
for w in $WHAT; do
  case $w in
    peertest) el "peer kernels on one device"; timeout 600 python -m pytest tests/test_peer_gpu.py -q -x 2>&1 | tail -4 ;;
    nccl)   el "weak, NCCL collectives (all-gather / reduce-scatter)"; run bench.py --gpus $N --no-cpu-baseline --no-e2e --halo-mode nccl --steps 100 --warmup 10 2> $O/nccl_n$N.err > $O/nccl_n$N.json; python tools/bench_summary.py < $O/nccl_n$N.json ;;
    strongnccl) el "strong, NCCL all-to-all"; run bench.py --gpus $N --no-cpu-baseline --no-e2e --scaling strong --halo-mode nccl --steps 100 --warmup 10 2> $O/strongnccl_n$N.err > $O/strongnccl_n$N.json; python tools/bench_summary.py < $O/strongnccl_n$N.json ;;
    sm)     el "weak, peer transport, all-gather by the store kernel"; run bench.py --gpus $N --no-cpu-baseline --no-e2e --peer-push sm --steps 100 --warmup 10 2> $O/sm_n$N.err > $O/sm_n$N.json; python tools/bench_summary.py < $O/sm_n$N.json ;;
    ce)     el "weak, peer transport, all-gather by copy engines"; run bench.py --gpus $N --no-cpu-baseline --no-e2e --peer-push ce --steps 100 --warmup 10 2> $O/ce_n$N.err > $O/ce_n$N.json; python tools/bench_summary.py < $O/ce_n$N.json ;;
    psparse) el "weak, peer transport, sparse-halo layout forced"; run bench.py --gpus $N --no-cpu-baseline --no-e2e --halo-mode peer_sparse --steps 100 --warmup 10 2> $O/psparse_n$N.err > $O/psparse_n$N.json; python tools/bench_summary.py < $O/psparse_n$N.json ;;
    push*)  v=${w#push}; el "weak, peer, all-gather store kernel with $v quarter-blocks per SM"; run bench.py --gpus $N --no-cpu-baseline --no-e2e --dev peer_push_blocks=$v --steps 100 --warmup 10 2> $O/push${v}_n$N.err > $O/push${v}_n$N.json; python tools/bench_summary.py < $O/push${v}_n$N.json | head -1 ;;
    check)  el "check"; run bench.py --gpus $N --check 2> $O/check_n$N.err | tee $O/check_n$N.json | cut -c1-900 ;;
    weak)   el "weak (auto exchange)"; run bench.py --gpus $N --no-cpu-baseline --steps 100 --warmup 10 2> $O/weak_n$N.err > $O/weak_n$N.json; python tools/bench_summary.py < $O/weak_n$N.json ;;
    a2a)    el "weak, all-to-all forced"; run bench.py --gpus $N --no-cpu-baseline --no-e2e --halo-mode alltoall --steps 100 --warmup 10 2> $O/a2a_n$N.err > $O/a2a_n$N.json; python tools/bench_summary.py < $O/a2a_n$N.json ;;
    strong) el "strong (ML-10M itself, nnz-balanced ranges)"; run bench.py --gpus $N --no-cpu-baseline --no-e2e --scaling strong --steps 100 --warmup 10 2> $O/strong_n$N.err > $O/strong_n$N.json; python tools/bench_summary.py < $O/strong_n$N.json ;;
  esac
done
el "done"; tail -2 $O/*_n$N.err 2>/dev/null | grep -v "^$" | tail -12
