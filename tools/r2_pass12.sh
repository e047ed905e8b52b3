set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_profile.log 2>&1; tail -15 gpurun_out/r2_pytest_profile.log
PP_KERNEL_PATH=1 timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_profile_kernels_equal_run_based_kernels --deselect tests/test_gpu_parity.py::test_profile_kernels_after_baum_welch > gpurun_out/r2_pytest_lane.log 2>&1; tail -3 gpurun_out/r2_pytest_lane.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_p2.log 2>&1; tail -1 gpurun_out/r2_bench_p2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e'])"
