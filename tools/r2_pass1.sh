# round 2, GPU pass 1: measured fp64 peaks, the k_fb_lane batch-size question, sanitizer evidence
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
./tools/fp64_peak 4000 > gpurun_out/r2_fp64_peak.jsonl 2>&1; cat gpurun_out/r2_fp64_peak.jsonl
for n in 200000 1000000; do
  python bench.py --workload baum-welch --bw-events $n --steps 3 --warmup 2 > gpurun_out/r2_bw_${n}.log 2>&1; tail -1 gpurun_out/r2_bw_${n}.log | cut -c1-900
done
# sanitizer: smoke batch (HMM kernels + segmenter) under memcheck, then racecheck
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_memcheck_smoke.log 2>&1; tail -5 gpurun_out/r2_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_racecheck_smoke.log 2>&1; tail -5 gpurun_out/r2_racecheck_smoke.log
