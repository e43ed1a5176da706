#!/usr/bin/env python3
"""Survivor statistics of candidate FAST phase-1 filters on the bench frame (seed 0, all 8 levels): how many pixels pass each necessary
condition, against the true corner fraction.  CPU only (numpy + the oracle pyramid).  See profiles/r02_fast_phase_attribution.md."""
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
import oracle_lib as O
from eorb_slam_b200 import synth
img=synth.make_frame(0)
orc=O.OrbOracle(1000,1.2,8,20,7,19,752,480)
orc.extract(img,(0,1000),True)
t=20
tot=0; totc=0; tots=0; tot2=0; tot3=0
ring=[(0,3),(1,3),(2,2),(3,1),(3,0),(3,-1),(2,-2),(1,-3),(0,-3),(-1,-3),(-2,-2),(-3,-1),(-3,0),(-3,1),(-2,2),(-1,3)]
for l in range(8):
    L=orc.level(l).astype(np.int16)
    h,w=L.shape
    c=L[3:-3,3:-3]
    def sh(dx,dy): return L[3+dy:h-3+dy,3+dx:w-3+dx]
    P=[sh(dx,dy) for dx,dy in ring]
    br=[p>c+t for p in P]; dk=[p<c-t for p in P]
    any_=[b|d for b,d in zip(br,dk)]
    compass=(any_[0]|any_[8])&(any_[4]|any_[12])
    # pair test no polarity all 8
    pair8=np.ones_like(compass)
    for k in range(8): pair8&=(any_[k]|any_[k+8])
    # polarity pair test (OpenCV)
    pb=np.ones_like(compass); pd=np.ones_like(compass)
    for k in range(8): pb&=(br[k]|br[k+8]); pd&=(dk[k]|dk[k+8])
    pol=pb|pd
    # 4 compass + 4 diag with polarity
    pb4=np.ones_like(compass); pd4=np.ones_like(compass)
    for k in (0,4,2,6): pb4&=(br[k]|br[k+8]); pd4&=(dk[k]|dk[k+8])
    pol4=pb4|pd4
    # exact corner
    def arc(fl):
        out=np.zeros_like(compass)
        for s in range(16):
            a=np.ones_like(compass)
            for j in range(9): a&=fl[(s+j)%16]
            out|=a
        return out
    corner=arc(br)|arc(dk)
    n=c.size
    print(l,n,"compass %.3f pair8 %.3f pol8 %.3f pol4 %.3f corner %.3f"%(compass.mean(),pair8.mean(),pol.mean(),pol4.mean(),corner.mean()))
    tot+=n; tots+=compass.sum(); totc+=corner.sum(); tot2+=pol.sum(); tot3+=pol4.sum()
print("all: compass %.3f pol8 %.3f pol4 %.3f corner %.3f"%(tots/tot,tot2/tot,tot3/tot,totc/tot))
