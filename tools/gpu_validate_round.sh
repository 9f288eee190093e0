# final validation of the round: whole GPU suite, smoke, the bench line, then the ncu launch list of the bench command
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02k_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02k_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02k_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r02k_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc $?" >> gpurun_out/r02k_bench.err
for c in vc60 rmvpe10; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r02k_cfg_$c.json 2> gpurun_out/r02k_cfg_$c.err; done
B200VC_CUDA_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02k_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-output-check > gpurun_out/r02k_bench_under_ncu.log 2>&1
gzip -f gpurun_out/r02k_launches.csv
ls -la gpurun_out | tail -8
