set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_effects_gpu.py tests/test_call_surface_gpu.py "tests/test_baseline_configs_gpu.py::test_cover_engine_stage_handoffs_30s" -m gpu -q -s -x > gpurun_out/r02h_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02h_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02h_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r02h_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc $?" >> gpurun_out/r02h_bench.err
timeout 300 python tools/debug_toy_mdx.py > gpurun_out/r02h_toy_mdx.log 2>&1
ls -la gpurun_out | tail -6
