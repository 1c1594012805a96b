#!/bin/bash
# Second validation call of the round: the two new GPU tests on their own, the full suite, the 512-px U-ViT training step
# (force_down_up_sample backward), and the ncu launch list of one base-256 train step with the final kernels.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_final2.log) 2>&1
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
summ() { grep -v "^$" "$1" | grep -E "^(FAILED|ERROR|E  |tests/.*(Error|assert)|[0-9]+ (passed|failed))|rel-L2" | cut -c1-400 | tail -${2:-25}; }
PDLTEST=tests/test_model_gpu.py::test_programmatic_dependent_launch_is_bit_identical
DUTEST=tests/test_uvit_v2_gpu.py::test_micro_uvit_v2_force_down_up_sample_training_gradients_vs_oracle
el "new tests, one process each"
timeout 300 python -m pytest $PDLTEST -q -s --tb=short -p no:cacheprovider > gpurun_out/r2_final2_pdltest.log 2>&1; echo "rc=$?"; summ gpurun_out/r2_final2_pdltest.log 40
timeout 300 python -m pytest $DUTEST -q -s --tb=short -p no:cacheprovider > gpurun_out/r2_final2_downup.log 2>&1; echo "rc=$?"; summ gpurun_out/r2_final2_downup.log 40
el "full suite"
timeout 700 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_final2_pytest.log 2>&1; echo "rc=$?"; summ gpurun_out/r2_final2_pytest.log
el "U-ViT 512-px training step (force_down_up_sample)"
timeout 300 python tools/bench_c4.py --model uvit512 --batch 16 --steps 3 --warmup 3 2> gpurun_out/r2_final2_uvit512.err | tee gpurun_out/r2_final2_uvit512.json; tail -2 gpurun_out/r2_final2_uvit512.err | cut -c1-300
el "launch list of one train step"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_step.csv 40 > gpurun_out/r02_launch_list_train_step_B256_final.txt; rm -f gpurun_out/launches_step.csv
head -16 gpurun_out/r02_launch_list_train_step_B256_final.txt
el "DONE"
