set -x
mkdir -p gpurun_out
SGG_L2_POLICY=1 python -m pytest tests/test_steps_gpu.py -m gpu -q -x > gpurun_out/pytest_r13.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r13.log
tail -3 gpurun_out/pytest_r13.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r13_a.json 2> gpurun_out/bench_r13_a.err; echo rc=$?
SGG_L2_POLICY=1 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r13_b.json 2> gpurun_out/bench_r13_b.err; echo rc=$?
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r13_c.json 2> gpurun_out/bench_r13_c.err; echo rc=$?
SGG_L2_POLICY=1 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r13_d.json 2> gpurun_out/bench_r13_d.err; echo rc=$?
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --timesteps 30 --vocab 5000 > gpurun_out/bench_r13_cfg3.json 2> gpurun_out/bench_r13_cfg3.err; echo rc=$?
