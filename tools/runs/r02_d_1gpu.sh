#!/bin/bash
# round 2, call D (1 GPU): parity tests of the rewritten two-sided kernel incl. the bodies' own order, reference render
# goldens, 4 vs 8 rows per lane, all configs
set -u
mkdir -p gpurun_out/r02d
O=gpurun_out/r02d
python tools/make_golden_render.py $O/render_golden.json > $O/render_golden.log 2>&1; echo "render golden rc=$?"; tail -3 $O/render_golden.log
cp $O/render_golden.json tests/golden/render_golden.json 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for rows in 4 8; do
  for cfg in disc1m cluster disc16k; do
    NBODY_B200_SYM_ROWS=$rows timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_${cfg}_rows$rows.json 2> $O/bench_${cfg}_rows$rows.err; echo "bench $cfg rows$rows rc=$?"
  done
done
timeout 300 python bench.py --config shipped --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_shipped.json 2> $O/bench_shipped.err
NBODY_B200_SYM_ROWS=8 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_sided or full_size or cell_sorted or rows_at" > $O/pytest_rows8.log 2>&1; echo "pytest rows8 rc=$?" >> $O/pytest_rows8.log
tail -3 $O/pytest_rows8.log
