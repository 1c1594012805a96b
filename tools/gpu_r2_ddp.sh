#!/bin/bash
# N-GPU checks: eager DDP vs the whole DDP step captured in a CUDA graph; config 4 under DDP.  NGPU from the environment.
N=${NGPU:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_ddp_n${N}.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@" 2>&1 | grep -E '^\{|Error|error' | cut -c1-900 | tail -3; }
echo "=== bench N=$N eager DDP"; run bench.py --gpus $N --steps 20 --warmup 3 --ddp-graph 0
echo "=== bench N=$N DDP step in one CUDA graph"; run bench.py --gpus $N --steps 20 --warmup 3 --ddp-graph 1
if [ -n "$C4" ]; then echo "=== config 4 (cc12m dims, B=64/GPU) N=$N"; run tools/bench_c4.py --steps 5 --warmup 3; fi
echo "=== DONE"
