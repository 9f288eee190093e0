set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt
python -m pytest tests -m gpu -q -s > gpurun_out/r02a_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02a_tests.log
B200VC_EXPERIMENTAL=1 python -m pytest tests -m gpu -q -s -k "fp16" > gpurun_out/r02a_tests_fp16.log 2>&1; echo "rc $?" >> gpurun_out/r02a_tests_fp16.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench_fp32.json 2> gpurun_out/r02a_bench_fp32.err
B200VC_MDX_FP16=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_fp16.json 2> gpurun_out/r02a_bench_fp16.err
B200VC_CUDA_GRAPHS=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-output-check > gpurun_out/r02a_bench_nograph.json 2> gpurun_out/r02a_bench_nograph.err
python tools/bench_tapgemm.py --only mdx --out gpurun_out/r02a_tapgemm_fp32.jsonl > gpurun_out/r02a_tapgemm_fp32.log 2>&1
python tools/bench_tapgemm.py --only mdx --fp16 --out gpurun_out/r02a_tapgemm_fp16.jsonl > gpurun_out/r02a_tapgemm_fp16.log 2>&1
for m in fp32 fp16; do
  F=""; [ $m = fp16 ] && F="--fp16"
  ncu --set full --clock-control none --import-source on -k regex:tapgemm_ws -c 1 -o gpurun_out/r02a_ws_c48_$m python tools/bench_tapgemm.py --iters 1 --warm 0 --only "mdx.l0 2d c48 B4" $F > gpurun_out/r02a_ncu_$m.log 2>&1
  ncu -i gpurun_out/r02a_ws_c48_$m.ncu-rep --page details > gpurun_out/r02a_ws_c48_${m}_details.txt 2>&1
  ncu -i gpurun_out/r02a_ws_c48_$m.ncu-rep --page raw --csv > gpurun_out/r02a_ws_c48_${m}_raw.csv 2>&1
  rm -f gpurun_out/r02a_ws_c48_$m.ncu-rep
done
ls -la gpurun_out | tail -20
