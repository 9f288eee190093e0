set -x
mkdir -p gpurun_out
python -m pytest tests/test_baseline_configs_gpu.py tests/test_rmvpe_gpu.py tests/test_call_surface_gpu.py tests/test_mdx_gpu.py tests/test_golden_gpu.py -m gpu -q -s > gpurun_out/r02c_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02c_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench_fp16.json 2> gpurun_out/r02c_bench_fp16.err
B200VC_MDX_FP16=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_tf32.json 2> gpurun_out/r02c_bench_tf32.err
B200VC_SYNTH_FP16=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_fp16both.json 2> gpurun_out/r02c_bench_fp16both.err
B200VC_SYNTH_FP16=1 B200VC_EXPERIMENTAL=1 python -m pytest tests/test_synth_gpu.py -m gpu -q -s -k fp16 > gpurun_out/r02c_tests_synth_fp16.log 2>&1
for c in rmvpe10 hubert30 vc60 mdx4min; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r02c_cfg_$c.json 2> gpurun_out/r02c_cfg_$c.err; done
ls -la gpurun_out | tail -14
