set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_r8.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r8.log
tail -5 gpurun_out/pytest_r8.log
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_r8.json 2> gpurun_out/bench_r8.err; echo bench_rc=$?
for c in 0 256 512 2048 8192; do
  python bench.py --workload sample --chunk $c --steps 6 --warmup 3 >> gpurun_out/bench_sample_r8.json 2>> gpurun_out/bench_sample_r8.err; echo rc=$?
done
python bench.py --workload sample --mode gumbel --steps 6 --warmup 3 >> gpurun_out/bench_sample_r8.json 2>> gpurun_out/bench_sample_r8.err
python tools/profile_iter.py > gpurun_out/plain_r8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_r8.csv python tools/profile_iter.py > gpurun_out/ncu_l_r8.log 2>&1
echo rc=$?
ncu --set full --clock-control none -k regex:'attn_|meanpool|adam_|lstm_' -s 60 -c 20 -o /tmp/prof_r8_hbm python tools/profile_iter.py --iters 1 > gpurun_out/ncu_h_r8.log 2>&1
echo rc=$?
ncu -i /tmp/prof_r8_hbm.ncu-rep --page raw --csv > gpurun_out/prof_r8_hbm.raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:gemm_kernel -s 185 -c 24 -o /tmp/prof_r8_gemm python tools/profile_iter.py --iters 1 > gpurun_out/ncu_g_r8.log 2>&1
echo rc=$?
ncu -i /tmp/prof_r8_gemm.ncu-rep --page raw --csv > gpurun_out/prof_r8_gemm.raw.csv 2>/dev/null
ls -la gpurun_out /tmp/*.ncu-rep
du -sh gpurun_out
