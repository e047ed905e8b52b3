# round 2: multi-GPU lines (N = $1), launched the way the driver launches them
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config2_${N}gpu.json 2> gpurun_out/r2_m1.err; tail -c 400 gpurun_out/r2_bench_config2_${N}gpu.json
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --scaling strong > gpurun_out/r2_bench_config2_${N}gpu_strong.json 2> gpurun_out/r2_m2.err; tail -c 400 gpurun_out/r2_bench_config2_${N}gpu_strong.json
$TR bench.py --gpus $N --workload baum-welch --bw-events 200000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_baumwelch_${N}gpu.json 2> gpurun_out/r2_m3.err; tail -c 600 gpurun_out/r2_bench_baumwelch_${N}gpu.json
$TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm_${N}gpu.json 2> gpurun_out/r2_m4.err; tail -c 300 gpurun_out/r2_bench_reference_arm_${N}gpu.json
python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -2
