set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_steps_gpu.py tests/test_sample_gpu.py -m gpu -q -x -k "fused or iteration or golden" > gpurun_out/pytest_r24.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r24.log
tail -3 gpurun_out/pytest_r24.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r24.json 2> gpurun_out/bench_r24.err; echo rc=$?
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r24b.json 2> gpurun_out/bench_r24b.err; echo rc=$?
