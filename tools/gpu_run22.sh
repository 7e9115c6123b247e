set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_steps_gpu.py -m gpu -q -x -k "fused or iteration" > gpurun_out/pytest_r22.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r22.log
tail -3 gpurun_out/pytest_r22.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r22.json 2> gpurun_out/bench_r22.err; echo rc=$?
