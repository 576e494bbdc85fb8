#!/bin/bash
# Round-2 opening sequence on the GPU box: everything that was written at the end of round 1 without GPU minutes left.
# Each step is bounded by its own timeout; run through gpurun, e.g.
#   gpurun --timeout 900 -- 'bash tools/round2_gpu.sh single > gpurun_out/round2_single.log 2>&1; tail -40 gpurun_out/round2_single.log'
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/round2_gpu.sh p2p > gpurun_out/round2_p2p.log 2>&1; tail -40 gpurun_out/round2_p2p.log'
set -u
case "${1:-single}" in
  single)
    timeout 600 python -m pytest tests -m gpu -x -q
    timeout 120 python tests/manual/edge_cases_gpu.py                 # promote what holds into tests/
    timeout 300 python bench.py --steps 3 --warmup 3
    ;;
  p2p)                                                                 # N >= 2: the peer-memory all-reduce against NCCL, then in the tree
    N=${2:-2}
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus "$N" --steps 3 --warmup 3
    TMC_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus "$N" --steps 3 --warmup 3
    ;;
esac
