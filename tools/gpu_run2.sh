# GPU job: parity tests of the tensor-core attention kernels, A/B bench (SIMT vs MMA attention), event-timed kernel table.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r2.log
tail -5 gpurun_out/pytest_r2.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo bench_rc=$?
SGG_ATTN_SIMT=1 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2_simt.json 2> gpurun_out/bench_r2_simt.err; echo rc=$?
python tools/profile_events.py > gpurun_out/events_r2.md 2> gpurun_out/events_r2.err; echo ev_rc=$?
ls -la gpurun_out
