set -x
GEMM_MICRO_ONLY=scores SGG_GEMM_MIN_SMEM=118000 python tools/gemm_micro.py > gpurun_out/gemm_micro_r18_pad.txt 2>&1
SGG_GEMM_MIN_SMEM=118000 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r18_pad.json 2> gpurun_out/bench_r18_pad.err; echo rc=$?
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r18_base.json 2> gpurun_out/bench_r18_base.err; echo rc=$?
