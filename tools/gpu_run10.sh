set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 300 $TR tests/dist/check_sharded.py > gpurun_out/check_sharded_n$N.log 2>&1; echo check_rc=$?
tail -3 gpurun_out/check_sharded_n$N.log
timeout 300 $TR bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_r10_n$N.json 2> gpurun_out/bench_r10_n$N.err; echo rc=$?
SGG_WA_SHARD=0 timeout 300 $TR bench.py --gpus $N --steps 50 --warmup 3 --no-e2e > gpurun_out/bench_r10_n${N}_replicated.json 2> gpurun_out/bench_r10_n${N}_replicated.err; echo rc=$?
timeout 300 $TR bench.py --gpus $N --workload sample --steps 6 --warmup 3 --chunk 8192 > gpurun_out/bench_r10_sample_n$N.json 2> gpurun_out/bench_r10_sample_n$N.err; echo rc=$?
cat gpurun_out/bench_r10_n$N.json | cut -c1-400
