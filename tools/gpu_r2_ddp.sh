#!/bin/bash
# N-GPU checks of bench.py's data-parallel path: NCCL SM reservation on / off, eager DDP vs the DDP step in one CUDA graph.
N=${NGPU:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_ddp_n${N}.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
run() { timeout ${RUN_TIMEOUT:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@" > gpurun_out/ddp_tmp.log 2>&1; echo "exit code $?"; grep -E '^\{' gpurun_out/ddp_tmp.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('   ', round(d['value']), d['unit'], round(d['ms_per_step'], 3), 'ms/step; e2e', round(d.get('e2e', {}).get('value', 0)), '| graph', d.get('config', {}).get('cuda_graph'), '| nccl sms', d.get('config', {}).get('nccl_sms_reserved'), '| ch', d.get('config', {}).get('nccl_max_channels'), '| diag', d.get('config', {}).get('DIAGNOSTIC_no_allreduce'))
"; grep -E "Error|error|Traceback" -A3 gpurun_out/ddp_tmp.log | grep -v "^--" | head -12; }
if [ -z "$VARIANTS" ]; then VARIANTS="eager graph nosync ch4 ch2"; fi
for v in $VARIANTS; do
  case $v in
    eager)  echo "=== N=$N eager DDP, NCCL default, full GEMM grids"; run bench.py --gpus $N --steps 20 --warmup 3 ;;
    graph)  echo "=== N=$N DDP step in one CUDA graph, NCCL default"; run bench.py --gpus $N --steps 20 --warmup 3 --ddp-graph 1 ;;
    nosync) echo "=== N=$N DIAGNOSTIC eager DDP without the all-reduce (launch overhead only)"; run bench.py --gpus $N --steps 20 --warmup 3 --ddp-no-sync ;;
    ch4)    echo "=== N=$N eager DDP, NCCL capped to 4 channels, full GEMM grids"; run bench.py --gpus $N --steps 20 --warmup 3 --nccl-channels 4 ;;
    ch2)    echo "=== N=$N eager DDP, NCCL capped to 2 channels, full GEMM grids"; run bench.py --gpus $N --steps 20 --warmup 3 --nccl-channels 2 ;;
    graphch4) echo "=== N=$N graph DDP, NCCL capped to 4 channels"; run bench.py --gpus $N --steps 20 --warmup 3 --ddp-graph 1 --nccl-channels 4 ;;
    sms4)   echo "=== N=$N eager DDP, 4 SMs reserved for NCCL"; run bench.py --gpus $N --steps 20 --warmup 3 --nccl-sms 4 ;;
    c4)     echo "=== config 4 (cc12m dims, B=64/GPU) N=$N"; run tools/bench_c4.py --steps 5 --warmup 3 ;;
  esac
done
echo "=== DONE"
