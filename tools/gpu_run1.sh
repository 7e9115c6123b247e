# Round-1 GPU job: parity tests, bench line, ncu launch list, ncu full-set captures of the dominant kernels.
# gpurun_out/ must stay under 64 MiB: the .ncu-rep files are converted to CSV on the box and dropped when large.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_r1.log
tail -3 gpurun_out/pytest_r1.log
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo bench_rc=$?
SGG_PDL=0 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1_nopdl.json 2> gpurun_out/bench_r1_nopdl.err; echo bench_nopdl_rc=$?
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1239 -c 830 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1; echo ncu1_rc=$?
ncu --set full --clock-control none --import-source on -k regex:"attn_fwd|attn_rev|adam|lstm_rev|meanpool" -c 14 -f -o gpurun_out/prof_r1_hbm $CMD > gpurun_out/ncu2.log 2>&1; echo ncu2_rc=$?
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 10 -f -o gpurun_out/prof_r1_gemm $CMD > gpurun_out/ncu3.log 2>&1; echo ncu3_rc=$?
for r in prof_r1_hbm prof_r1_gemm; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page source --csv > gpurun_out/$r.source.csv 2>/dev/null
  gzip -f gpurun_out/$r.source.csv
  sz=$(stat -c %s gpurun_out/$r.ncu-rep); if [ "$sz" -gt 20000000 ]; then rm gpurun_out/$r.ncu-rep; fi
done
du -sh gpurun_out; ls -la gpurun_out
