set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r9.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r9.log
tail -15 gpurun_out/pytest_r9.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r9_n1.json 2> gpurun_out/bench_r9_n1.err; echo rc=$?
SGG_PROJ_PREFETCH=0 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r9_n1_noprefetch.json 2> gpurun_out/bench_r9_n1_noprefetch.err; echo rc=$?
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500"
timeout 300 $TR bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_r9_n2.json 2> gpurun_out/bench_r9_n2.err; echo rc=$?
SGG_WA_SHARD=0 timeout 300 $TR bench.py --gpus 2 --steps 50 --warmup 3 --no-e2e > gpurun_out/bench_r9_n2_replicated.json 2> gpurun_out/bench_r9_n2_replicated.err; echo rc=$?
tail -3 gpurun_out/*.err
