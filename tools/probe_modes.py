import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/vision-based-spatio-temporal-analysis_b200'); sys.path.insert(0,'/root/repo/tools')
import sweep_variants as sv
from bevipm import rig
wl = rig.WORKLOADS[sys.argv[1] if len(sys.argv)>1 else "c2"]
for v in [int(x) for x in (sys.argv[2] if len(sys.argv)>2 else "31,40,41").split(",")]:
    for mode in ("sum","mean"):
        r = sv.time_variant(wl, v, mode=mode)
        print(wl.name, v, mode, {k: (round(x,4) if isinstance(x,float) else x) for k,x in r.items() if k in ("ms","error")})
