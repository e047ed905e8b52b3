set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_profile.log 2>&1; tail -15 gpurun_out/r2_pytest_profile.log
PP_KERNEL_PATH=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_lane.log 2>&1; tail -3 gpurun_out/r2_pytest_lane.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_profile.log 2>&1; tail -1 gpurun_out/r2_bench_profile.log | cut -c1-1500
./tools/lat_bench > gpurun_out/r2_lat.log 2>&1; tail -6 gpurun_out/r2_lat.log
