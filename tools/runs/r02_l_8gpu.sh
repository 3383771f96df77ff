#!/bin/bash
# round 2, call L (8 GPUs): the world-4 tests, the two-sided flow at 8 GPUs against one GPU, the bench line at 1 / 4 / 8 GPUs,
# BASELINE configs[4] (N = 4 194 304 two-galaxy) on 8 GPUs
set -u
mkdir -p gpurun_out/r02l
O=gpurun_out/r02l
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 10 --warmup 5 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench8 rc=$?"
timeout 400 $TR --nproc-per-node 4 bench.py --gpus 4 --steps 10 --warmup 5 > $O/bench_4gpu.json 2> $O/bench_4gpu.err; echo "bench4 rc=$?"
timeout 400 $TR --nproc-per-node 2 bench.py --gpus 2 --steps 10 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench1 rc=$?"
timeout 600 $TR --nproc-per-node 8 bench.py --gpus 8 --config galaxy --steps 3 --warmup 3 > $O/bench_galaxy_8gpu.json 2> $O/bench_galaxy_8gpu.err; echo "galaxy rc=$?"
timeout 300 $TR --nproc-per-node 8 tools/mgpu_sym_check.py 1048576 3 > $O/mgpu_sym_8_1m.log 2> $O/mgpu_sym_8_1m.err; echo "sym 8 rc=$?"
timeout 300 $TR --nproc-per-node 4 tools/mgpu_sym_check.py 131072 3 > $O/mgpu_sym_4_131072.log 2> $O/mgpu_sym_4_131072.err; echo "sym 4 rc=$?"
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "4]" > $O/pytest_world4.log 2>&1; echo "pytest rc=$?" >> $O/pytest_world4.log
tail -4 $O/pytest_world4.log
grep -h replicas $O/*.log | cut -c1-300
