set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500"
timeout 200 $TR bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/bench_final_n8.json 2> gpurun_out/bench_final_n8.err; echo rc=$?
