set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; tail -30 gpurun_out/r2_pytest_all.log
