set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_r4.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r4.log
tail -5 gpurun_out/pytest_r4.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r4.json 2> gpurun_out/bench_r4.err; echo bench_rc=$?
python tools/profile_events.py > gpurun_out/events_r4.md 2> gpurun_out/events_r4.err; echo ev_rc=$?
