set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_r5.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r5.log
tail -5 gpurun_out/pytest_r5.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r5.json 2> gpurun_out/bench_r5.err; echo bench_rc=$?
