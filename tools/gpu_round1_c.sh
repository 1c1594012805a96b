#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1_c.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_sampling_vqgan_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | cut -c1-400 | tail -150
echo "=== DONE"
