#!/bin/bash
# N-GPU checks of bench.py's data-parallel path: NCCL SM reservation on / off, eager DDP vs the DDP step in one CUDA graph.
N=${NGPU:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_ddp_n${N}.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@" > gpurun_out/ddp_tmp.log 2>&1; echo "exit code $?"; grep -E '^\{' gpurun_out/ddp_tmp.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('   ', round(d['value']), d['unit'], round(d['ms_per_step'], 3), 'ms/step; e2e', round(d['e2e']['value']), '| graph', d['config'].get('cuda_graph'), '| nccl sms', d['config'].get('nccl_sms_reserved'))
"; grep -E "Error|error|Traceback" -A3 gpurun_out/ddp_tmp.log | grep -v "^--" | head -12; }
echo "=== N=$N eager DDP, NCCL default, full GEMM grids"; run bench.py --gpus $N --steps 20 --warmup 3 --nccl-sms 0
echo "=== N=$N eager DDP, 4 SMs reserved for NCCL"; run bench.py --gpus $N --steps 20 --warmup 3 --nccl-sms 4
echo "=== N=$N DDP step in one CUDA graph, 4 SMs reserved"; run bench.py --gpus $N --steps 20 --warmup 3 --nccl-sms 4 --ddp-graph 1
echo "=== N=$N eager DDP, 8 SMs reserved"; run bench.py --gpus $N --steps 20 --warmup 3 --nccl-sms 8
if [ -n "$C4" ]; then echo "=== config 4 (cc12m dims, B=64/GPU) N=$N"; run tools/bench_c4.py --steps 5 --warmup 3; fi
echo "=== DONE"
