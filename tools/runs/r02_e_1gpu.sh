#!/bin/bash
# round 2, call E (1 GPU): render goldens (fixed), full GPU suite incl. the deeper parity tests and the merge on every path,
# free-running divergence data, small-n sweep two-sided vs one-sided, ncu captures of the force kernel at three sizes
set -u
mkdir -p gpurun_out/r02e
O=gpurun_out/r02e
python tools/make_golden_render.py $O/render_golden.json > $O/render_golden.log 2>&1; echo "render golden rc=$?"; tail -2 $O/render_golden.log
cp $O/render_golden.json tests/golden/render_golden.json 2>/dev/null
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python tools/free_running_divergence.py 60 > $O/free_running.jsonl 2> $O/free_running.err; echo "free running rc=$?"
for N in 8192 16384 24576 32768 40000; do
  for minn in 6144 1000000; do
    NBODY_B200_SYM_MIN_N=$minn timeout 300 python bench.py --config disc16k --n $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-whole-run > $O/sweep_${N}_symmin$minn.json 2> $O/sweep_${N}_symmin$minn.err; echo "sweep $N $minn rc=$?"
  done
done
python tools/prof_step.py disc1m 3 > $O/plain_1m.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:force_sym -s 2 -c 1 -o $O/force_sym_1m python tools/prof_step.py disc1m 3 > $O/ncu_1m.log 2>&1
echo "ncu 1m rc=$?"
python tools/prof_step.py disc16k 4 > $O/plain_16k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:force_sym -s 3 -c 1 -o $O/force_sym_16k python tools/prof_step.py disc16k 4 > $O/ncu_16k.log 2>&1
echo "ncu 16k rc=$?"
python tools/prof_step.py cluster 6 > $O/plain_cluster.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:force_sym -s 5 -c 1 -o $O/force_sym_cluster python tools/prof_step.py cluster 6 > $O/ncu_cluster.log 2>&1
echo "ncu cluster rc=$?"
python tools/prof_step.py disc16k 6 > $O/plain_16k_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file $O/launches_16k.csv python tools/prof_step.py disc16k 6 > $O/ncu_l16k.log 2>&1
echo "launch list rc=$?"
ls -la $O | head -50
