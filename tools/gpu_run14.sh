set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r14.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r14.log
tail -3 gpurun_out/pytest_r14.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r14_a.json 2> gpurun_out/bench_r14_a.err; echo rc=$?
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r14_b.json 2> gpurun_out/bench_r14_b.err; echo rc=$?
