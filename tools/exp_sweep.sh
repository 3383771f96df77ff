for e in NOWAIT NOFOLD NOTEST; do echo "== $e"; NBODY_B200_LIB=build/exp_$e/libnbody_b200.so timeout 200 python tools/variant_sweep.py 131072 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('v%d ok=%s frac=%.3f regs=%d'%(d['variant'],d['parity_ok'],d['frac_roofline_force'],d['regs']))
"; done
