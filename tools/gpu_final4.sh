set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29500"
timeout 200 $TR bench.py --gpus 4 --steps 50 --warmup 3 > gpurun_out/bench_final_n4.json 2> gpurun_out/bench_final_n4.err; echo rc=$?
