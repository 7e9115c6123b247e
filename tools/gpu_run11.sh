set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500"
timeout 300 $TR tests/dist/check_sharded.py > gpurun_out/check_sharded_r11.log 2>&1; echo check_rc=$?
grep -E "check_sharded ok|AssertionError" gpurun_out/check_sharded_r11.log | head -5
python -m pytest tests/test_dropin_gpu.py tests/test_sample_gpu.py "tests/test_steps_gpu.py" -m gpu -q -x -k "recall or attention or sample or gumbel or greedy or trainer or iteration" > gpurun_out/pytest_r11.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r11.log
tail -4 gpurun_out/pytest_r11.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r11_n1.json 2> gpurun_out/bench_r11_n1.err; echo rc=$?
SGG_WA_PITCH=dense python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r11_n1_dense.json 2> gpurun_out/bench_r11_n1_dense.err; echo rc=$?
