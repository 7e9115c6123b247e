set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_steps_gpu.py tests/test_dropin_gpu.py -m gpu -q -x > gpurun_out/pytest_r20.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r20.log
tail -12 gpurun_out/pytest_r20.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r20_fused.json 2> gpurun_out/bench_r20_fused.err; echo rc=$?
SGG_FUSED_ADAM_PROJ=0 timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r20_sep.json 2> gpurun_out/bench_r20_sep.err; echo rc=$?
