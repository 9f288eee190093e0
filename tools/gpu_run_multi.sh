# usage: bash tools/gpu_run_multi.sh N   (inside gpurun --gpus N)
set -x
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "$N" -le 2 ]; then timeout 600 $TR tools/check_strong_scaling.py > gpurun_out/r02_strong_check_n${N}_small.log 2>&1; echo "rc $?" >> gpurun_out/r02_strong_check_n${N}_small.log; fi
timeout 900 $TR tools/check_strong_scaling.py --full > gpurun_out/r02_strong_check_n${N}_full.log 2>&1; echo "rc $?" >> gpurun_out/r02_strong_check_n${N}_full.log
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
grep -E "rank|rc " gpurun_out/r02_strong_check_n${N}_*.log | tail -20
