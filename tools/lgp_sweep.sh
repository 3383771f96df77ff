#!/bin/bash
# Work-unit size sweep (tile parts, NBODY_B200_LG_PARTS) at several n; the calibration of plan_fill's thresholds.
#   gpurun -- 'bash tools/lgp_sweep.sh "16384 32768 65536"'   (default sizes below; unsorted order so that only the unit size varies)
SIZES=${1:-"16384 32768 65536 131072 262144 524288"}
for n in $SIZES; do for lg in 0 1 2 3; do echo -n "n=$n lgP=$lg: "; NBODY_B200_LG_PARTS=$lg python - <<PY
import sys
sys.path.insert(0,'.')
import numpy as np
import __graft_entry__ as G
nb=G.load_package()
n=$n
R=1e5*np.sqrt(n/16384.0); field=int(R)
block0=nb.generate(nb.SCENARIO_DISC,n,extent=R,field_w=field,field_h=field)
for var in (0,1):
    sim=nb.Simulation(n,field_w=field,field_h=field,coverage=nb.COVERAGE_FULL,flags=nb.flag_variant(var)|nb.FLAG_NO_SORT)
    sim.upload(block0,n); sim.step(3); sim.sync()
    s0=sim.stats(); tot,frc=sim.step_timed(5,force=True); s1=sim.stats()
    pairs=s1['pairs']-s0['pairs']
    print('v%d force_ms=%.4f frac=%.3f exact=%d fast=%d |'%(var,frc/5,pairs*20/(frc*1e-3)/74.45e12,s1['exact_chunks']-s0['exact_chunks'],s1['fast_chunks']-s0['fast_chunks']),end=' ')
    sim.close()
print()
PY
done; done
