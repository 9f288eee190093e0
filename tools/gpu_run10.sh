set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_effects_gpu.py tests/test_pipeline_gpu.py "tests/test_baseline_configs_gpu.py::test_cover_engine_stage_handoffs_30s" -m gpu -q -s > gpurun_out/r02j_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02j_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc $?" >> gpurun_out/r02j_bench.err
ls -la gpurun_out | tail -4
