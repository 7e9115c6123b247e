set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final2.log 2>&1; echo smoke_rc=$?
tail -2 gpurun_out/smoke_final2.log
python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_final2.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_final2.log
tail -3 gpurun_out/pytest_final2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500"
timeout 300 $TR bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_final_n2.json 2> gpurun_out/bench_final_n2.err; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_n2_ref.json 2> gpurun_out/bench_final_n2_ref.err; echo rc=$?
