#!/usr/bin/env bash
# The GPU-side validation sequence of a round, one gpurun call per line (run from the repo root; each call is
# charged separately, ncu only after the same command has exited 0 without it).
#   tools/gpu_validate.sh tests | bench | ncu | launches | mgpu2 | scale N
set -euo pipefail
GPURUN=/usr/local/graft/bin/gpurun
case "${1:-}" in
tests)
    $GPURUN --timeout 900 -- 'timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/gpu_tests.log; cat gpurun_out/gpu_tests.log; timeout 60 python -c "import __graft_entry__ as G; G.smoke()" 2>&1 | tail -2' ;;
bench)
    $GPURUN --timeout 400 -- 'timeout 300 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; head -c 400 gpurun_out/bench_1gpu.json' ;;
ncu)
    $GPURUN --timeout 600 -- 'timeout 120 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/pre_ncu.json 2> gpurun_out/pre_ncu.err && timeout 400 ncu --set full --clock-control none --import-source on -k regex:force_sym_kernel -s 3 -c 1 -o gpurun_out/force_sym_1m -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_sym.log 2>&1; tail -3 gpurun_out/ncu_sym.log'
    echo 'then: python tools/ncu_summary.py gpurun_out/force_sym_1m.ncu-rep "<command>" > profiles/rNN_force_1m_ncu.json' ;;
launches)
    $GPURUN --timeout 600 -- 'timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/pre_ncu.json 2> gpurun_out/pre_ncu.err && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1; wc -l gpurun_out/launches.csv' ;;
mgpu2)
    $GPURUN --gpus 2 --timeout 600 -- 'timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/mgpu_tests.log; cat gpurun_out/mgpu_tests.log' ;;
scale)
    N="${2:?scale needs the GPU count}"
    $GPURUN --gpus "$N" --timeout 400 -- "timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; head -c 300 gpurun_out/bench_${N}gpu.json" ;;
*)
    sed -n 2,5p "$0"; exit 2 ;;
esac
