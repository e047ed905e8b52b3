set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_profile.log 2>&1; tail -5 gpurun_out/r2_pytest_profile.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_p2.log 2>&1; tail -1 gpurun_out/r2_bench_p2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms_per_step'])"
PROTPORE_B200_LIB=$PWD/protpore_b200/libprotpore_b200_alt1.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_p1.log 2>&1; tail -1 gpurun_out/r2_bench_p1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms_per_step'])"
ncu --set full --clock-control none --import-source on -k regex:"k_profile" -s 2 -c 2 -o gpurun_out/r2_profile_a -f python bench.py --events 30000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
