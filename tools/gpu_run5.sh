set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02e_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02e_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
B200VC_SYNTH_FP16=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_synthfp16.json 2> gpurun_out/r02e_bench_synthfp16.err
timeout 600 python tools/stage_breakdown.py --fine > gpurun_out/r02e_breakdown.json 2> gpurun_out/r02e_breakdown.err
ls -la gpurun_out | tail -8
