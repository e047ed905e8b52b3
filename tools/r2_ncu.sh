set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_profile" -s 2 -c 2 -o gpurun_out/r2_profile_c -f python bench.py --events 30000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
