#!/bin/bash
# round 2, call G (1 GPU): warp-level kernel, unroll sweep
set -u
mkdir -p gpurun_out/r02g
O=gpurun_out/r02g
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "own_order or full_coverage or disc_scenario or conserving" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for U in 2 4 8 32; do
  for N in 4096 16384 32768; do
    NBODY_B200_SYMW_UNROLL=$U timeout 300 python bench.py --config disc16k --n $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-whole-run > $O/sweep_${N}_u$U.json 2> $O/sweep_${N}_u$U.err; echo "sweep $N u$U rc=$?"
  done
done
python tools/prof_step.py disc16k 4 > $O/plain_16k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:force_symw -s 3 -c 1 -o $O/force_symw_16k python tools/prof_step.py disc16k 4 > $O/ncu_16k.log 2>&1
echo "ncu 16k rc=$?"
