#!/bin/bash
# launch lists (ncu gpu__time_duration) of the secondary hot paths: decode loop, tokenizer encode (both precisions), decode, U-ViT step
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_paths.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
ll() { PROFILE_PATH=$1 PROFILE_CONV=$2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$3.csv python tools/profile_paths.py > /dev/null 2>&1
  python tools/summarize_launches.py gpurun_out/launches_$3.csv 40 > gpurun_out/r02_launch_list_$3.txt; rm -f gpurun_out/launches_$3.csv; echo "=== $3"; head -${4:-24} gpurun_out/r02_launch_list_$3.txt; }
ll decode bf16x3 generate2_B64
ll encode bf16x3 vq_encode_bf16x3_B64
ll encode bf16 vq_encode_bf16_B64
ll vqdecode bf16x3 vq_decode_B64
ll uvit bf16x3 uvit_step_B16
echo "=== DONE"
