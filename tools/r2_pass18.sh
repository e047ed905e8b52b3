set -x
mkdir -p gpurun_out
python bench.py --workload config4 --c4-events 1024 --steps 2 --warmup 3 > gpurun_out/r2_bench_config4.log 2>&1; tail -1 gpurun_out/r2_bench_config4.log | cut -c1-1800
python bench.py --workload config5 --c5-events 200000 --steps 2 --warmup 3 > gpurun_out/r2_bench_config5.log 2>&1; tail -1 gpurun_out/r2_bench_config5.log | cut -c1-1800
python bench.py --workload baum-welch --bw-events 200000 --steps 2 --warmup 3 > gpurun_out/r2_bench_bw.log 2>&1; tail -1 gpurun_out/r2_bench_bw.log | cut -c1-1800
