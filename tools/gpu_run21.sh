set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sample_gpu.py tests/test_steps_gpu.py -m gpu -q -x -k "sample or greedy or gumbel or chunk or fused or argmax" > gpurun_out/pytest_r21.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r21.log
tail -5 gpurun_out/pytest_r21.log
for c in 0 2048 8192; do
  timeout 300 python bench.py --workload sample --chunk $c --steps 6 --warmup 3 >> gpurun_out/bench_sample_r21.json 2>> gpurun_out/bench_sample_r21.err; echo rc=$?
done
SGG_SAMPLE_STREAMS=1 timeout 300 python bench.py --workload sample --chunk 4096 --steps 6 --warmup 3 >> gpurun_out/bench_sample_r21.json 2>> gpurun_out/bench_sample_r21.err; echo rc=$?
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r21.json 2> gpurun_out/bench_r21.err; echo rc=$?
