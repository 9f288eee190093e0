# ncu evidence for profiles/: launch list of the bench command + one --set full capture of the dominant family's representative launch
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_synth_gpu.py tests/test_variants_gpu.py tests/test_hubert_gpu.py tests/test_pipeline_gpu.py tests/test_golden_gpu.py "tests/test_baseline_configs_gpu.py::test_cfg3_vc_pipeline_60s_with_ivf2237_index" -m gpu -q -s > gpurun_out/r02f_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02f_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
B200VC_CUDA_GRAPHS=0 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-output-check > gpurun_out/r02f_bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r02f_launches.csv > gpurun_out/r02f_launch_summary.txt 2>&1
for shape in "mdx.l2 2d c144 B4" "mdx.l1 2d c96 B4" "mdx.l0 2d c48 B4"; do
  tag=$(echo $shape | tr ' .' '__')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm -c 1 -o gpurun_out/r02f_$tag python tools/bench_tapgemm.py --iters 1 --warm 0 --only "$shape" --fp16 > gpurun_out/r02f_ncu_$tag.log 2>&1
  ncu -i gpurun_out/r02f_$tag.ncu-rep --page details > gpurun_out/r02f_${tag}_details.txt 2>&1
  ncu -i gpurun_out/r02f_$tag.ncu-rep --page raw --csv > gpurun_out/r02f_${tag}_raw.csv 2>&1
  rm -f gpurun_out/r02f_$tag.ncu-rep
done
gzip -f gpurun_out/r02f_launches.csv
ls -la gpurun_out | tail -12
