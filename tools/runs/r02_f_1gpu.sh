#!/bin/bash
# round 2, call F (1 GPU): the warp-level two-sided kernel for small n: full GPU suite, then the small-n sweep
# (warp-level vs CTA-level vs one-sided) and every config through bench.py
set -u
mkdir -p gpurun_out/r02f
O=gpurun_out/r02f
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -6 $O/pytest.log
for N in 2048 4096 8192 16384 24576 32768 40000; do
  for mode in warp cta one; do
    case $mode in
      warp) export NBODY_B200_SYM_SMALL=2 NBODY_B200_SYM_MIN_N=1024;;
      cta) export NBODY_B200_SYM_SMALL=1 NBODY_B200_SYM_MIN_N=1024;;
      one) export NBODY_B200_SYM_SMALL=2 NBODY_B200_SYM_MIN_N=100000000;;
    esac
    timeout 300 python bench.py --config disc16k --n $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-whole-run > $O/sweep_${N}_$mode.json 2> $O/sweep_${N}_$mode.err; echo "sweep $N $mode rc=$?"
  done
done
unset NBODY_B200_SYM_SMALL NBODY_B200_SYM_MIN_N
for cfg in disc16k shipped cluster disc1m; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "bench $cfg rc=$?"
done
python tools/prof_step.py disc16k 4 > $O/plain_16k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:force_symw -s 3 -c 1 -o $O/force_symw_16k python tools/prof_step.py disc16k 4 > $O/ncu_16k.log 2>&1
echo "ncu 16k rc=$?"
