set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_steps_gpu.py tests/test_dropin_gpu.py -m gpu -q -x -k "fit or trainer or recall or graph" > gpurun_out/pytest_r25.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r25.log
tail -3 gpurun_out/pytest_r25.log
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r25.json 2> gpurun_out/bench_r25.err; echo rc=$?
