set -x
mkdir -p gpurun_out
python tools/profile_iter.py > gpurun_out/plain_r7.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_r7.csv python tools/profile_iter.py > gpurun_out/ncu_l_r7.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:'attn_|meanpool|adam_|lstm_|pack_hl|colsum' -s 60 -c 36 -o gpurun_out/prof_r7_hbm python tools/profile_iter.py --iters 1 > gpurun_out/ncu_h_r7.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 185 -c 40 -o gpurun_out/prof_r7_gemm python tools/profile_iter.py --iters 1 > gpurun_out/ncu_g_r7.log 2>&1
echo rc=$?
ls -la gpurun_out
