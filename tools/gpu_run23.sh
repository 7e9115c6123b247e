set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_steps_gpu.py -m gpu -q -x -k "disc_step or gen_step or iteration" > gpurun_out/pytest_r23.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r23.log
tail -3 gpurun_out/pytest_r23.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r23.json 2> gpurun_out/bench_r23.err; echo rc=$?
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r23b.json 2> gpurun_out/bench_r23b.err; echo rc=$?
