import sys, json
sys.path.insert(0,'.')
import numpy as np
import __graft_entry__ as G
nb=G.load_package()
for n in (8192, 16384, 24576, 32768, 49152, 65536, 98304, 262144):
    R=1e5*np.sqrt(n/16384.0); field=int(R)
    block0=nb.generate(nb.SCENARIO_DISC,n,extent=R,field_w=field,field_h=field)
    for smn, fl in ((0, 0), (1024, nb.FLAG_ONE_SIDED), (1024, 0)):
        flags = nb.FLAG_NO_SORT if smn == 0 else fl
        sim=nb.Simulation(n,field_w=field,field_h=field,coverage=nb.COVERAGE_FULL,flags=flags,sort_min_n=smn)
        sim.upload(block0,n); sim.step(3); sim.sync()
        s0=sim.stats(); tot,_=sim.step_timed(10,force=False); s1=sim.stats()
        pairs=s1['pairs']-s0['pairs']
        print(json.dumps({"n":n,"sorted":smn>0,"two_sided":bool(s1["pair_halving"]),"graph_step_ms":tot/10,"frac_step":pairs*20/(tot*1e-3)/74.45e12}),flush=True)
        sim.close()
