#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1_c.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 900 python -m pytest tests/test_sampling_vqgan_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "conv or vqgan" 2>&1 | grep -v "^$" | cut -c1-400 | tail -20
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "conditional" -s 2>&1 | grep -v "^$" | cut -c1-600 | tail -12
VQ_BATCH=32 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_vqgan.csv python tools/profile_vqgan.py > /dev/null
python tools/summarize_launches.py gpurun_out/launches_vqgan.csv 30
echo "=== AUX"; timeout 600 python tools/bench_aux.py vqgan
echo "=== C4 skipped"
echo "=== DONE"
