set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_effects_gpu.py tests/test_call_surface_gpu.py tests/test_pipeline_gpu.py "tests/test_baseline_configs_gpu.py::test_cover_engine_stage_handoffs_30s" -m gpu -q -s > gpurun_out/r02i_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02i_tests.log
timeout 600 python tools/check_f0_overlap.py > gpurun_out/r02i_f0_overlap.log 2>&1; echo "rc $?" >> gpurun_out/r02i_f0_overlap.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc $?" >> gpurun_out/r02i_bench.err
B200VC_F0_OVERLAP=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_overlap.json 2> gpurun_out/r02i_bench_overlap.err; echo "bench rc $?" >> gpurun_out/r02i_bench_overlap.err
ls -la gpurun_out | tail -6
