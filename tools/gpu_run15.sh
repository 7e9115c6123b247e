set -x
mkdir -p gpurun_out
python -m pytest tests/test_steps_gpu.py tests/test_dropin_gpu.py -m gpu -q -x > gpurun_out/pytest_r15.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r15.log
tail -3 gpurun_out/pytest_r15.log
for pen in 0 3000 1000 10000; do
  SGG_SPLIT_PENALTY=$pen python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r15_p$pen.json 2> gpurun_out/bench_r15_p$pen.err; echo rc=$?
done
