#!/bin/bash
# round 2, call N (1 GPU): the warp-level kernel at large n against the CTA-level kernel
set -u
mkdir -p gpurun_out/r02n
O=gpurun_out/r02n
for N in 262144 524288; do
  for mode in cta warp; do
    if [ $mode = warp ]; then export NBODY_B200_SYMW_MAX_N=100000000; else export NBODY_B200_SYMW_MAX_N=40960; fi
    timeout 120 python bench.py --config disc1m --n $N --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-whole-run > $O/maxn_${N}_$mode.json 2> $O/maxn_${N}_$mode.err; echo "maxn $N $mode rc=$?"
  done
done
NBODY_B200_SYMW_MAX_N=100000000 timeout 200 python bench.py --config disc1m --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-whole-run > $O/disc1m_warp.json 2> $O/disc1m_warp.err; echo "1m warp rc=$?"
