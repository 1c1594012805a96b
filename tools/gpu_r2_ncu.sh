#!/bin/bash
# round-2 evidence: ncu --set full of every named kernel (one launch each) + the launch list of one train step
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_ncu.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 1500 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/r2_kernels python tools/profile_kernels.py > gpurun_out/r2_kernels_run.log 2>&1
tail -2 gpurun_out/r2_kernels_run.log
python tools/make_profile_summary.py gpurun_out/r2_kernels.ncu-rep gpurun_out/r02_ncu_kernels.txt > /dev/null
wc -l gpurun_out/r02_ncu_kernels.txt
ls -la gpurun_out/r2_kernels.ncu-rep
if [ $(stat -c %s gpurun_out/r2_kernels.ncu-rep) -gt 40000000 ]; then rm gpurun_out/r2_kernels.ncu-rep; echo "(report too large to bring back; summary kept)"; fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2.csv python tools/profile_step.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2.csv 60 > gpurun_out/r02_launch_list_train_step_B256.txt
head -45 gpurun_out/r02_launch_list_train_step_B256.txt
echo "=== DONE"
