set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_r6.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r6.log
tail -5 gpurun_out/pytest_r6.log
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_r6.json 2> gpurun_out/bench_r6.err; echo bench_rc=$?
python tools/profile_events.py > gpurun_out/events_r6.md 2> gpurun_out/events_r6.err; echo ev_rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r6.json 2> gpurun_out/bench_ref_r6.err; echo ref_rc=$?
