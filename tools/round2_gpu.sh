#!/bin/bash
# Multi-GPU checks of round 2 (run through gpurun --gpus N):
#   bash tools/round2_gpu.sh check N    the N-rank tree equals the 1-rank tree through every entry path; peer-memory all-reduce vs NCCL
#   bash tools/round2_gpu.sh bench N    bench.py at N ranks with NCCL and with TMC_P2P=1 (c3), hashes checked against the single-GPU tree
set -u
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${2:-2}
case "${1:-check}" in
  check)
    timeout 600 $TR --nproc-per-node "$N" --master-port 29521 tools/mgpu_check.py tiny small c1
    timeout 300 $TR --nproc-per-node "$N" --master-port 29511 tools/p2p_check.py
    ;;
  bench)
    W=${3:-c3}
    timeout 900 $TR --nproc-per-node "$N" --master-port 29512 bench.py --gpus "$N" --steps 3 --warmup 2 --workload "$W" > gpurun_out/r2_bench_${W}_n${N}_nccl.json 2> gpurun_out/r2_bench_${W}_n${N}_nccl.err
    TMC_P2P=1 timeout 900 $TR --nproc-per-node "$N" --master-port 29513 bench.py --gpus "$N" --steps 3 --warmup 2 --workload "$W" > gpurun_out/r2_bench_${W}_n${N}_p2p.json 2> gpurun_out/r2_bench_${W}_n${N}_p2p.err
    for f in nccl p2p; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_${W}_n${N}_$f.json"))
    print("$W N=$N $f value %.1f e2e %.1f steps %d sweeps %d hash %s %s" % (d["value"], d["e2e"]["value"], d["tree"]["lanczos_steps"], d["tree"]["sweeps"], d["tree"]["hash"][:10], d["tree"]["hash_check"]))
except Exception as e:
    print("$W N=$N $f FAILED", e); print(open("gpurun_out/r2_bench_${W}_n${N}_$f.err").read()[-800:])
PY
    done
    ;;
esac
