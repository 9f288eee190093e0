set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rmvpe_gpu.py tests/test_call_surface_gpu.py tests/test_crepe_gpu.py -m gpu -q -s > gpurun_out/r02g_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02g_tests.log
timeout 600 python tools/debug_toy_mdx.py > gpurun_out/r02g_toy_mdx.log 2>&1
timeout 600 python bench.py --config crepe60 --steps 3 --warmup 3 > gpurun_out/r02g_cfg_crepe60.json 2> gpurun_out/r02g_cfg_crepe60.err
for v in 1 3; do B200VC_GRU=$v timeout 300 python bench.py --config rmvpe10 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_cfg_rmvpe10_gru$v.json 2> gpurun_out/r02g_cfg_rmvpe10_gru$v.err; done
ls -la gpurun_out | tail -6
