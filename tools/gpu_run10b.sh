set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_effects_gpu.py -m gpu -q -s > gpurun_out/r02j2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02j2_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02j2_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r02j2_smoke.log
