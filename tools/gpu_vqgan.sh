#!/bin/bash
# VQGAN / tokenizer loop: tests of the conv + VQ + pipeline paths, then the config-3 aux bench on both conv routes
mkdir -p gpurun_out
exec > >(tee gpurun_out/vqgan.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 900 python -m pytest tests/test_sampling_vqgan_gpu.py tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "vq or conv or vqgan or pipeline" 2>&1 | grep -v "^$" | cut -c1-400 | tail -40
echo "=== AUX tc"; VQ_BATCH=${VQ_BATCH:-64} timeout 600 python tools/bench_aux.py vqgan
echo "=== AUX simt"; MUSE_B200_CONV=simt VQ_BATCH=16 timeout 600 python tools/bench_aux.py vqgan
echo "=== DONE"
