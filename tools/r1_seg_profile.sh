set -x
python bench.py --workload segment --steps 5 --warmup 3 > gpurun_out/bench_segment_final.log 2>&1; tail -1 gpurun_out/bench_segment_final.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_segment_launches.csv \
    python bench.py --workload segment --seg-events 4000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_seg_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_statsplit|k_cumsum|k_segments" -s 3 -c 3 -o gpurun_out/prof_r1_statsplit3 -f \
    python bench.py --workload segment --seg-events 4000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_seg3.log 2>&1
ncu -i gpurun_out/prof_r1_statsplit3.ncu-rep --page details --csv > gpurun_out/r1_statsplit_details.csv 2>/dev/null
