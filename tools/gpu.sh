#!/bin/bash
# One parameterised runner for the B200 box (replaces the per-call scratch scripts of round 1).
#   gpurun --timeout 900 -- 'bash tools/gpu.sh test bench'            # GPU tests, then the default bench line
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu.sh dist:8 bench:8'
# Stages (run in the order given; each writes gpurun_out/<tag>_<stage>.* and a one-line status to stdout):
#   test[:expr]     python -m pytest tests -m gpu -q -x [-k expr]
#   smoke           __graft_entry__.smoke()
#   bench[:N]       bench.py (N ranks under torchrun when N > 1); extra flags from $BENCH_FLAGS
#   ref             bench.py --impl reference --steps 3 --warmup 1
#   dist:N          tests/dist/check_sharded.py under torchrun with N ranks; extra flags from $DIST_FLAGS
#   launches        ncu launch list (gpu__time_duration) of tools/profile_iter.py
#   launches_bench  ncu launch list of the bench command itself
#   full:<regex>    ncu --set full capture of kernels matching <regex> in $NCU_SCRIPT (default tools/profile_iter.py; -c $NCU_COUNT, default 6)
#   events          tools/profile_events.py (SGG_TIMING=1 per-launch CUDA-event timing)
#   py:<file>       python <file> (stdout/stderr to gpurun_out)
# TAG (env, default r2) prefixes the output files.
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29500"
NCU="ncu --clock-control none"
for stage in "$@"; do
  name=${stage%%:*}; arg=""; [[ "$stage" == *:* ]] && arg=${stage#*:}
  out=gpurun_out/${TAG}_${name}${arg:+_${arg//[^A-Za-z0-9]/_}}
  case "$name" in
    test)
      timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q -x ${arg:+-k "$arg"} > $out.log 2>&1; rc=$?
      tail -5 $out.log ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out.log 2>&1; rc=$?; tail -2 $out.log ;;
    bench)
      n=${arg:-1}
      if [ "$n" -gt 1 ]; then
        timeout 900 $TR --nproc-per-node $n bench.py --gpus $n ${BENCH_FLAGS:-} > $out.json 2> $out.err; rc=$?
      else
        timeout 900 python bench.py ${BENCH_FLAGS:-} > $out.json 2> $out.err; rc=$?
      fi
      tail -c 600 $out.json ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out.json 2> $out.err; rc=$?; tail -c 400 $out.json ;;
    dist)
      timeout 900 $TR --nproc-per-node $arg tests/dist/check_sharded.py ${DIST_FLAGS:-} > $out.log 2>&1; rc=$?
      grep -E "check_sharded|Error|error" $out.log | tail -8 ;;
    launches)
      timeout 900 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $out.csv python tools/profile_iter.py > $out.log 2>&1; rc=$? ;;
    launches_bench)
      timeout 900 $NCU --metrics gpu__time_duration.sum -c 8000 --csv --log-file $out.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $out.log 2>&1; rc=$? ;;
    full)
      timeout 900 $NCU --set full --import-source on ${NCU_EXTRA:-} -k "regex:$arg" -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-6} -o $out -f python ${NCU_SCRIPT:-tools/profile_iter.py} > $out.log 2>&1; rc=$?
      [ -f $out.ncu-rep ] && ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null ;;
    events)
      SGG_TIMING=1 SGG_PDL=0 timeout 600 python tools/profile_events.py > $out.md 2> $out.err; rc=$? ;;
    py)
      timeout ${PY_TIMEOUT:-900} python $arg > gpurun_out/${TAG}_$(basename $arg .py).log 2>&1; rc=$?
      tail -15 gpurun_out/${TAG}_$(basename $arg .py).log ;;
    *) echo "unknown stage $stage"; rc=99 ;;
  esac
  echo "== $stage rc=$rc"
done
