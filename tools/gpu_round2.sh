#!/bin/bash
# Round-1 second GPU session: batched multi-view tests first (all failures shown), then the rest of the GPU suite,
# the config sweep with the batched cube-face path, and a short bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_views.py -q --tb=short -p no:cacheprovider > gpurun_out/pytest_views.log 2>&1
echo "views rc=$?" >> gpurun_out/pytest_views.log
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider --ignore=tests/test_gpu_views.py > gpurun_out/pytest_gpu.log 2>&1
echo "gpu rc=$?" >> gpurun_out/pytest_gpu.log
if [ "$1" != "testsonly" ]; then
timeout 900 python tests/tools/run_configs.py --skip5 > gpurun_out/configs.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
fi
tail -40 gpurun_out/pytest_views.log; tail -15 gpurun_out/pytest_gpu.log; cut -c1-1500 gpurun_out/configs.log | tail -20; cut -c1-600 gpurun_out/bench.json
