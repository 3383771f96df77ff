#!/bin/bash
# round 2, call S (1 GPU): self-aware pre-tested loop for the own chunks of the warp-level kernel
set -u
mkdir -p gpurun_out/r02s
O=gpurun_out/r02s
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "warp_level or own_order or full_coverage or disc_scenario or conserving or collapsing or cluster_131072 or plummer" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 100 python bench.py --config disc16k --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_disc16k.json 2> $O/bench_disc16k.err; echo "bench 16k rc=$?"
timeout 100 python bench.py --config cluster --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cluster.json 2> $O/bench_cluster.err; echo "bench cluster rc=$?"
