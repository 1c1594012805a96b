#!/bin/bash
# First thing to run on the next B200 call: the GPU tests written after round 2's GPU budget was spent (use_conv_in_out,
# the unmodified train_muse.py), WITHOUT their non-strict xfail marks, then the validated suite.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_pending.sh'
mkdir -p gpurun_out
exec > >(tee gpurun_out/pending.log) 2>&1
timeout 400 python -m pytest tests/test_zz_pending_gpu.py -m gpu --runxfail -q --tb=short -p no:cacheprovider -s; echo "pending rc=$?"
timeout 700 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pending_full_suite.log 2>&1; echo "suite rc=$?"
tail -3 gpurun_out/pending_full_suite.log
echo "=== DONE"
