set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_final3.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_final3.log
tail -3 gpurun_out/pytest_final3.log
python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo bench_rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final3.log 2>&1; echo smoke_rc=$?
