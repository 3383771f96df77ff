#!/bin/bash
# round 2, call O (1 GPU): final validation -- the whole GPU suite, every BASELINE config through both bench arms, ncu of the
# final force kernel at n = 1M
set -u
mkdir -p gpurun_out/r02o
O=gpurun_out/r02o
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for cfg in disc1m disc16k shipped cluster; do
  timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "bench $cfg rc=$?"
  timeout 300 python bench.py --impl reference --config $cfg --steps 3 --warmup 1 > $O/ref_$cfg.json 2> $O/ref_$cfg.err; echo "ref $cfg rc=$?"
done
timeout 300 python bench.py --config galaxy --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_galaxy.json 2> $O/bench_galaxy.err; echo "bench galaxy rc=$?"
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_default_20.json 2> $O/bench_default_20.err; echo "bench 20 rc=$?"
python tools/prof_step.py disc1m 3 > $O/plain_1m.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:force_sym_kernel -s 2 -c 1 -o $O/force_sym_1m python tools/prof_step.py disc1m 3 > $O/ncu_1m.log 2>&1
echo "ncu 1m rc=$?"
python tools/prof_step.py cluster 4 > $O/plain_cluster.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 30 --csv --log-file $O/launches_cluster.csv python tools/prof_step.py cluster 4 > $O/ncu_lc.log 2>&1
echo "launch list rc=$?"
