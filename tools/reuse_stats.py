#!/usr/bin/env python
"""How much tap re-use a walking order offers on the synthetic rig (CPU only, uses the oracle's coordinates).

Part 1: for cell tiles TH x TW, unique 2x2 blocks and unique texels per seen cell-view (what a register /
shared-memory cache of that footprint could save over 4 loads per cell-view).
Part 2: walking a BEV row, how consecutive seen cells of one view relate: same block, shifted by one texel
in x, in y, or a new block.  These numbers decided the design of csrc/ipm_run.cuh."""
import sys, numpy as np, torch
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'vision-based-spatio-temporal-analysis_b200'))
from oracle import ipm_oracle as O
from bevipm import rig
for name in ("c2","c3"):
    wl = rig.WORKLOADS[name]
    K,Rt = rig.look_at_rig(7,0)
    xs,ys = rig.ground_axes(*wl.bev_hw, wl.bounds)
    ix,iy = O.coords(K[None].numpy(),Rt[None].numpy(),xs.numpy(),ys.numpy(),wl.feat_hw,wl.img_size)
    ix=ix.reshape(7,*wl.bev_hw); iy=iy.reshape(7,*wl.bev_hw)
    Hf,Wf=wl.feat_hw
    x0=np.floor(ix).astype(np.int64); y0=np.floor(iy).astype(np.int64)
    vis = (x0>=-1)&(x0<Wf)&(y0>=-1)&(y0<Hf)
    print(name,"visible cell-views",vis.sum(), "of", vis.size)
    blk = y0*100000+x0
    def stats(TH,TW):
        Hb,Wb=wl.bev_hw
        tot_cv=0; tot_blocks=0; tot_texels=0; runs_row=0
        for v in range(7):
            for i0 in range(0,Hb,TH):
                for j0 in range(0,Wb,TW):
                    m=vis[v,i0:i0+TH,j0:j0+TW]
                    if not m.any(): continue
                    b=blk[v,i0:i0+TH,j0:j0+TW][m]
                    tot_cv+=m.sum()
                    ub=np.unique(b); tot_blocks+=len(ub)
                    # unique texels
                    yy=ub//100000; xx=ub-yy*100000
                    # handle negative properly
                    yy=y0[v,i0:i0+TH,j0:j0+TW][m]; xx=x0[v,i0:i0+TH,j0:j0+TW][m]
                    t=set()
                    for dy in (0,1):
                        for dx in (0,1):
                            t.update(zip((yy+dy).tolist(),(xx+dx).tolist()))
                    tot_texels+=len(t)
        return tot_cv,tot_blocks,tot_texels
    for TH,TW in ((1,8),(1,16),(1,32),(2,4),(2,8),(4,4),(4,8),(8,8),(8,1),(16,1)):
        cv,bl,tx=stats(TH,TW)
        print(f"  tile {TH}x{TW}: cell-views {cv}, unique blocks {bl} (x{cv/bl:.2f}), unique texels {tx}: loads/cv {tx/cv:.2f} (vs 4)")


# ---- part 2: transitions along a row ----
for name in ("c2","c3"):
    wl = rig.WORKLOADS[name]
    K,Rt = rig.look_at_rig(7,0)
    xs,ys = rig.ground_axes(*wl.bev_hw, wl.bounds)
    ix,iy = O.coords(K[None].numpy(),Rt[None].numpy(),xs.numpy(),ys.numpy(),wl.feat_hw,wl.img_size)
    ix=ix.reshape(7,*wl.bev_hw); iy=iy.reshape(7,*wl.bev_hw)
    Hf,Wf=wl.feat_hw
    x0=np.floor(ix).astype(np.int64); y0=np.floor(iy).astype(np.int64)
    vis = (x0>=-1)&(x0<Wf)&(y0>=-1)&(y0<Hf)
    for TW in (8,16):
        tot=dict(cv=0,same=0,xs=0,ys=0,full=0)
        per_view=[]
        for v in range(7):
            d=dict(cv=0,same=0,xs=0,ys=0,full=0)
            Wb=wl.bev_hw[1]
            c=np.arange(Wb)%TW
            dx=np.diff(x0[v],axis=1,prepend=x0[v][:,:1]); dy=np.diff(y0[v],axis=1,prepend=y0[v][:,:1])
            pv=np.concatenate([np.zeros((vis.shape[1],1),bool),vis[v][:,:-1]],axis=1)
            first=(c==0)[None,:]|~pv
            s=vis[v]
            same=s&~first&(dx==0)&(dy==0)
            xsft=s&~first&(np.abs(dx)==1)&(dy==0)
            ysft=s&~first&(dx==0)&(np.abs(dy)==1)
            full=s&~same&~xsft&~ysft
            d['cv']=s.sum(); d['same']=same.sum(); d['xs']=xsft.sum(); d['ys']=ysft.sum(); d['full']=full.sum()
            per_view.append(d)
            for k in tot: tot[k]+=d[k]
        print(name,"TW",TW,{k:round(v/tot['cv'],3) for k,v in tot.items()}, "texel loads/cv: blockrun %.2f, +xshift %.2f, +x,y shift %.2f"%(
            4*(tot['xs']+tot['ys']+tot['full'])/tot['cv'], (2*tot['xs']+4*tot['ys']+4*tot['full'])/tot['cv'], (2*tot['xs']+2*tot['ys']+4*tot['full'])/tot['cv']))
        for v,d in enumerate(per_view): print("   view",v,{k:round(x/max(d['cv'],1),2) for k,x in d.items() if k!='cv'}, d['cv'])
