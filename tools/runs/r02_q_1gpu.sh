#!/bin/bash
# round 2, call Q (1 GPU): last validation -- smoke(), the whole GPU suite, the 16k line after the launch pruning
set -u
mkdir -p gpurun_out/r02q
O=gpurun_out/r02q
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 100 python bench.py --config disc16k --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_disc16k.json 2> $O/bench_disc16k.err; echo "bench 16k rc=$?"
