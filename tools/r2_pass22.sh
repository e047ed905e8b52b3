set -x
for st in 1 0; do
for n in 200000 1000000; do
  if [ $st = 1 ]; then export PP_FB_STRIDE=1; else unset PP_FB_STRIDE; fi
  python bench.py --workload baum-welch --bw-events $n --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stride_sorted' if '$st'=='1' else 'stride_mixed', $n, d['e_step_kernel_ms_per_step'], d['e_step_gcups'], d['non_kernel_ms_per_step'])"
done; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "forward_backward or baum or tiny" 2>&1 | tail -3
