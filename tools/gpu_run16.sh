set -x
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r16_base.json 2> gpurun_out/bench_r16_base.err; echo rc=$?
SGG_PROJ_PREFETCH=0 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r16_nopre.json 2> gpurun_out/bench_r16_nopre.err; echo rc=$?
SGG_SIDE_STREAM=0 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r16_noside.json 2> gpurun_out/bench_r16_noside.err; echo rc=$?
SGG_PDL=0 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r16_nopdl.json 2> gpurun_out/bench_r16_nopdl.err; echo rc=$?
