set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02b_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02b_tests.log
B200VC_EXPERIMENTAL=1 python -m pytest tests -m gpu -q -s -k "fp16" > gpurun_out/r02b_tests_fp16.log 2>&1; echo "rc $?" >> gpurun_out/r02b_tests_fp16.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02b_bench_fp32.json 2> gpurun_out/r02b_bench_fp32.err
B200VC_MDX_FP16=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_fp16.json 2> gpurun_out/r02b_bench_fp16.err
B200VC_MDX_FP16=1 B200VC_SYNTH_FP16=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_fp16both.json 2> gpurun_out/r02b_bench_fp16both.err
B200VC_MDX_FP16=1 python tools/stage_breakdown.py --fine > gpurun_out/r02b_breakdown_fp16.json 2> gpurun_out/r02b_breakdown_fp16.err
ls -la gpurun_out | tail -12
