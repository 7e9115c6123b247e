set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500"
timeout 300 $TR tests/dist/check_sharded.py > gpurun_out/check_sharded_r12.log 2>&1; echo check_rc=$?
grep -E "check_sharded|AssertionError" gpurun_out/check_sharded_r12.log | head -20
python -m pytest tests/test_gemm_gpu.py tests/test_sample_gpu.py -m gpu -q -x > gpurun_out/pytest_r12.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r12.log
tail -3 gpurun_out/pytest_r12.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r12_n1.json 2> gpurun_out/bench_r12_n1.err; echo rc=$?
