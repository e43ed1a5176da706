import os, sys, math
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from eorb_slam_b200 import api, synth
img = synth.make_frame(0)
ex = api.ORBextractor(api.ORBxParams())
orc = O.OrbOracle()
ex(img); orc.extract(img)
E=19
for l in (0,7):
    lv = orc.level(l); H,W = lv.shape
    minB=E-3; maxBX=W-E+3; maxBY=H-E+3
    width=np.float32(maxBX-minB); height=np.float32(maxBY-minB)
    nC=int(width/np.float32(30)); nR=int(height/np.float32(30))
    wC=int(math.ceil(width/np.float32(nC))); hC=int(math.ceil(height/np.float32(nR)))
    gx,gy,gs = ex.debug_candidates(l)
    print("level",l,"grid",nC,nR,wC,hC,"gpu total",len(gx))
    shown=0
    for i in range(nR):
        for j in range(nC):
            iniY=minB+i*hC; maxY=min(iniY+hC+6,maxBY); iniX=minB+j*wC; maxX=min(iniX+wC+6,maxBX)
            if iniY>=maxBY-3 or iniX>=maxBX-3: continue
            roi=np.ascontiguousarray(lv[iniY:maxY,iniX:maxX])
            xs,ys,sc=O.fast(roi,20,True)
            if len(xs)==0: xs,ys,sc=O.fast(roi,7,True)
            ox=xs+j*wC; oy=ys+i*hC
            # gpu candidates of this cell: x_rel in [j*wC+3, j*wC + (maxX-iniX) -3)
            sel=(gx>=j*wC+3)&(gx<j*wC+(maxX-iniX)-3)&(gy>=i*hC+3)&(gy<i*hC+(maxY-iniY)-3)
            g=sorted(zip(gy[sel].tolist(),gx[sel].tolist(),gs[sel].tolist())); o=sorted(zip(oy.tolist(),ox.tolist(),sc.tolist()))
            if g!=o and shown<6:
                shown+=1
                print(" cell",i,j,"x0",iniX,"aoff",iniX&3,"w,h",maxX-iniX,maxY-iniY,"gpu",len(g),"cpu",len(o))
                print("   gpu",g[:8]); print("   cpu",o[:8])
    print(" mismatching cells shown",shown)
