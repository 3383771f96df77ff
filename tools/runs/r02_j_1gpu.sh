#!/bin/bash
# round 2, call J (1 GPU): warp-level kernel with own items first + queue
set -u
mkdir -p gpurun_out/r02j
O=gpurun_out/r02j
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "warp_level or own_order or full_coverage or conserving" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
for N in 8192 12288 16384 24576 32768 40000; do
  for mode in default one; do
    case $mode in
      default) F=0;;
      one) F=36;;
    esac
    NBODY_B200_SYM_MIN_N=8192 timeout 300 python bench.py --config disc16k --n $N --flags $F --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-whole-run > $O/sweep_${N}_$mode.json 2> $O/sweep_${N}_$mode.err; echo "sweep $N $mode rc=$?"
  done
done
timeout 300 python bench.py --config disc16k --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_disc16k.json 2> $O/bench_disc16k.err
python tools/prof_step.py disc16k 4 > $O/plain_16k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:force_symw -s 3 -c 1 -o $O/force_symw_16k python tools/prof_step.py disc16k 4 > $O/ncu_16k.log 2>&1
echo "ncu 16k rc=$?"
