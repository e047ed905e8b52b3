set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_all.log 2>&1; tail -8 gpurun_out/r2_pytest_all.log
PP_KERNEL_PATH=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --deselect tests/test_gpu_parity.py::test_profile_kernels_equal_run_based_kernels --deselect tests/test_gpu_parity.py::test_profile_kernels_after_baum_welch > gpurun_out/r2_pytest_lane.log 2>&1; tail -3 gpurun_out/r2_pytest_lane.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_p2.log 2>&1; tail -1 gpurun_out/r2_bench_p2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e'])"
timeout 600 python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_c5.log 2>&1; tail -1 gpurun_out/r2_bench_c5.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
