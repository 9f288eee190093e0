set -x
mkdir -p gpurun_out
timeout 600 python tools/debug_determinism.py > gpurun_out/r02d_determinism.log 2>&1
timeout 900 python -m pytest tests/test_rmvpe_gpu.py tests/test_call_surface_gpu.py "tests/test_baseline_configs_gpu.py::test_cover_engine_stage_handoffs_30s" -m gpu -q -s > gpurun_out/r02d_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02d_tests.log
timeout 300 python bench.py --config rmvpe10 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_cfg_rmvpe10.json 2> gpurun_out/r02d_cfg_rmvpe10.err
B200VC_GRU_V1=1 timeout 300 python bench.py --config rmvpe10 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_cfg_rmvpe10_gruv1.json 2> gpurun_out/r02d_cfg_rmvpe10_gruv1.err
timeout 900 python -m pytest tests/test_crepe_gpu.py -m gpu -q -s > gpurun_out/r02d_tests_crepe.log 2>&1; echo "rc $?" >> gpurun_out/r02d_tests_crepe.log
ls -la gpurun_out | tail -6
