set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_r3.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r3.log
tail -5 gpurun_out/pytest_r3.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r3.json 2> gpurun_out/bench_r3.err; echo bench_rc=$?
SGG_SIDE_STREAM=0 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r3_noside.json 2> gpurun_out/bench_r3_noside.err; echo bench_rc=$?
python tools/gemm_micro.py > gpurun_out/gemm_micro_r3.txt 2> gpurun_out/gemm_micro_r3.err; echo gm_rc=$?
SGG_PDL=0 python tools/gemm_micro.py > gpurun_out/gemm_micro_r3_nopdl.txt 2>&1; echo gm_rc=$?
