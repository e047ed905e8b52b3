# round 2, the evidence pass: all gpu tests on both kernel families, smoke, the bench lines, then ncu captures
# (each only after its command has passed without ncu)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
PP_KERNEL_PATH=1 python -m pytest tests/test_gpu_parity.py -m gpu -q --deselect tests/test_gpu_parity.py::test_profile_kernels_equal_run_based_kernels --deselect tests/test_gpu_parity.py::test_profile_kernels_after_baum_welch --deselect tests/test_gpu_parity.py::test_profile_forward_backward_equals_run_based_and_exact_kernels > gpurun_out/r2_pytest_gpu_lane_family.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu_lane_family.log
PP_GUARD=1 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_guarded.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu_guarded.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py --steps 10 --warmup 5 > gpurun_out/r2_bench_config2_1gpu.json 2> gpurun_out/r2_bench_config2.err; tail -c 600 gpurun_out/r2_bench_config2_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>&1; tail -c 300 gpurun_out/r2_bench_reference_arm.json
python bench.py --workload config4 --steps 2 --warmup 3 > gpurun_out/r2_bench_config4_1gpu.json 2>gpurun_out/r2_c4.err; tail -c 300 gpurun_out/r2_bench_config4_1gpu.json
python bench.py --workload config5 --steps 3 --warmup 3 > gpurun_out/r2_bench_config5_1gpu.json 2>gpurun_out/r2_c5.err; tail -c 300 gpurun_out/r2_bench_config5_1gpu.json
python bench.py --workload baum-welch --bw-events 200000 --steps 3 --warmup 3 > gpurun_out/r2_bench_baumwelch_1gpu.json 2>gpurun_out/r2_bw.err; tail -c 300 gpurun_out/r2_bench_baumwelch_1gpu.json
python bench.py --workload segment --steps 5 --warmup 3 > gpurun_out/r2_bench_segment_1gpu.json 2>gpurun_out/r2_seg.err; tail -c 300 gpurun_out/r2_bench_segment_1gpu.json
python bench.py --workload align --steps 5 --warmup 3 > gpurun_out/r2_bench_align_1gpu.json 2>gpurun_out/r2_al.err; tail -c 300 gpurun_out/r2_bench_align_1gpu.json
python bench.py --impl reference --workload align > gpurun_out/r2_bench_reference_arm_align.json 2>&1; tail -c 200 gpurun_out/r2_bench_reference_arm_align.json
# launch list of the contract command (short form) and full captures of the two sweeps
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --events 200000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_profile" -s 6 -c 2 -o gpurun_out/r2_profile_kernels -f \
    python bench.py --events 30000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_profile_fb -c 1 -o gpurun_out/r2_profile_fb -f \
    python bench.py --workload baum-welch --bw-events 20000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pfb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_segment_align -s 2 -c 1 -o gpurun_out/r2_segalign -f python tools/align_time.py > gpurun_out/ncu_sa.log 2>&1
# DRAM traffic of ONE Viterbi sweep at the full config-2 size
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_profile_viterbi -s 3 -c 1 --csv \
    --log-file gpurun_out/r2_viterbi_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_vt.log 2>&1
tail -3 gpurun_out/r2_viterbi_traffic.csv
