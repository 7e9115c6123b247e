set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench_rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo ref_rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo smoke_rc=$?
python bench.py --workload sample --steps 6 --warmup 3 > gpurun_out/bench_final_sample.json 2> gpurun_out/bench_final_sample.err; echo rc=$?
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --timesteps 30 --vocab 5000 > gpurun_out/bench_final_cfg3.json 2> gpurun_out/bench_final_cfg3.err; echo rc=$?
python tools/profile_iter.py > gpurun_out/plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv python tools/profile_iter.py > gpurun_out/ncu_l_final.log 2>&1
echo rc=$?
ncu --set full --clock-control none -k regex:'attn_|adam_' -s 40 -c 10 -o /tmp/prof_final_hbm python tools/profile_iter.py --iters 1 > gpurun_out/ncu_h_final.log 2>&1
echo rc=$?
ncu -i /tmp/prof_final_hbm.ncu-rep --page raw --csv > gpurun_out/prof_final_hbm.raw.csv 2>/dev/null
du -sh gpurun_out
