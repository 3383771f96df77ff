#!/bin/bash
# round 2, call H (1 GPU): stale sort (order carried over compactions) + warp-level kernel on the sorted order
set -u
mkdir -p gpurun_out/r02h
O=gpurun_out/r02h
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -6 $O/pytest.log
for N in 12288 16384 24576 32768 40000 49152 65536; do
  for mode in default nosort one; do
    case $mode in
      default) F=0;;
      nosort) F=4;;
      one) F=36;;
    esac
    timeout 300 python bench.py --config disc16k --n $N --flags $F --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-whole-run > $O/sweep_${N}_$mode.json 2> $O/sweep_${N}_$mode.err; echo "sweep $N $mode rc=$?"
  done
done
for cfg in disc16k shipped cluster disc1m; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "bench $cfg rc=$?"
done
python tools/prof_step.py disc16k 6 > $O/plain_16k.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file $O/launches_16k.csv python tools/prof_step.py disc16k 6 > $O/ncu_l16k.log 2>&1
echo "launch list rc=$?"
