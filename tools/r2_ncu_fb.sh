set -x
mkdir -p gpurun_out
export PP_FB_STRIDE=1
for n in 200000 1000000; do
ncu --set full --clock-control none -k regex:"k_fb_lane" -s 1 -c 1 -o gpurun_out/r2_fb_$n -f python bench.py --workload baum-welch --bw-events $n --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fb_$n.log 2>&1; tail -2 gpurun_out/ncu_fb_$n.log
done
