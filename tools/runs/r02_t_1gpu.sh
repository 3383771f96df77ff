#!/bin/bash
# round 2, call T (1 GPU): the final build -- smoke() and the whole GPU suite once more
set -u
mkdir -p gpurun_out/r02t
O=gpurun_out/r02t
timeout 60 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
