set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r1.log
tail -3 gpurun_out/pytest_r1.log
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo bench_rc=$?
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1239 -c 830 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1; echo ncu1_rc=$?
ncu --set full --clock-control none --import-source on -s 1652 -c 100 -f -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1; echo ncu2_rc=$?
ls -la gpurun_out
