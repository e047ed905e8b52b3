# one GPU pass: all gpu tests, smoke, bench lines, then ncu captures (each only after its command passed without ncu)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --workload segment --steps 5 --warmup 3 > gpurun_out/bench_segment.log 2>&1; tail -1 gpurun_out/bench_segment.log | cut -c1-400
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_score.log 2>&1; tail -1 gpurun_out/bench_score.log | cut -c1-300
# launch list + full capture of the segmenter
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_segment_launches.csv \
    python bench.py --workload segment --seg-events 4000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_seg_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_statsplit -s 1 -c 1 -o gpurun_out/prof_r1_statsplit -f \
    python bench.py --workload segment --seg-events 4000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_seg.log 2>&1
# DRAM traffic of the Viterbi sweep at the FULL config-2 size (one launch)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_viterbi_lane -s 1 -c 1 --csv \
    --log-file gpurun_out/r1_viterbi_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_vt.log 2>&1
tail -3 gpurun_out/r1_viterbi_traffic.csv
