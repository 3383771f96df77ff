#!/bin/bash
# round 2, call M (1 GPU): A/B of the "direction first" sub-step form (scalar-mass accumulates), and how far up the
# warp-level kernel stays ahead of the CTA-level one
set -u
mkdir -p gpurun_out/r02m
O=gpurun_out/r02m
for lib in base uv; do
  if [ $lib = uv ]; then export NBODY_B200_LIB=$PWD/build/uv/libnbody_b200_uv.so; else unset NBODY_B200_LIB; fi
  for cfg in disc1m cluster; do
    timeout 200 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-whole-run > $O/ab_${cfg}_$lib.json 2> $O/ab_${cfg}_$lib.err; echo "ab $cfg $lib rc=$?"
  done
done
unset NBODY_B200_LIB
for N in 65536 131072; do
  for mode in cta warp; do
    if [ $mode = warp ]; then export NBODY_B200_SYMW_MAX_N=100000000; else unset NBODY_B200_SYMW_MAX_N; fi
    timeout 120 python bench.py --config disc16k --n $N --batch 5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-whole-run > $O/maxn_${N}_$mode.json 2> $O/maxn_${N}_$mode.err; echo "maxn $N $mode rc=$?"
  done
done
