set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_profile.log 2>&1; tail -5 gpurun_out/r2_pytest_profile.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_p2.log 2>&1; tail -1 gpurun_out/r2_bench_p2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms_per_step'])"
python tools/profile_phase_timing.py 4736
python tools/profile_phase_timing.py 200000
