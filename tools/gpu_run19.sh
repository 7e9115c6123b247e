set -x
export GEMM_MICRO_ONLY=dWa
python tools/gemm_micro.py > gpurun_out/dwa_base.txt 2>&1
SGG_NO_WIDE_B=1 python tools/gemm_micro.py > gpurun_out/dwa_nowide.txt 2>&1
SGG_NO_WIDE_B=1 SGG_GEMM_MAX_STAGES=1 python tools/gemm_micro.py > gpurun_out/dwa_nowide_s1.txt 2>&1
SGG_NO_WIDE_B=1 SGG_GEMM_MAX_STAGES=2 python tools/gemm_micro.py > gpurun_out/dwa_nowide_s2.txt 2>&1
SGG_GEMM_MAX_STAGES=1 python tools/gemm_micro.py > gpurun_out/dwa_wide_s1.txt 2>&1
unset GEMM_MICRO_ONLY
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r19.json 2> gpurun_out/bench_r19.err; echo rc=$?
