#!/bin/bash
# round 2, call I (1 GPU): what bounds the warp-level kernel at n = 16384? round robin vs queue, chunks per item
set -u
mkdir -p gpurun_out/r02i
O=gpurun_out/r02i
for N in 16384 32768; do
 for Q in 0 1; do
  for R in 0 2 4; do
    NBODY_B200_SYMW_QUEUE=$Q NBODY_B200_SYMW_RUN=$R timeout 300 python bench.py --config disc16k --n $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-whole-run > $O/sweep_${N}_q${Q}_r$R.json 2> $O/sweep_${N}_q${Q}_r$R.err; echo "sweep $N q$Q r$R rc=$?"
  done
 done
done
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "warp_level or own_order" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
python tools/prof_step.py disc16k 4 > $O/plain_16k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:force_symw -s 3 -c 1 -o $O/force_symw_16k python tools/prof_step.py disc16k 4 > $O/ncu_16k.log 2>&1
echo "ncu 16k rc=$?"
