import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/oracle"); sys.path.insert(0, ROOT + "/dynamic-visual-slam_b200/python")
import numpy as np
import orbx, c_oracle as co
w, h = 1280, 720
g = co.synth_gray(7, 0, w, h)
ex = orbx.ORBextractor(max_width=w, max_height=h)
orc = co.COracle()
ref = orc.extract(g, trace=True)
kps, desc = ex(g)
for l in range(8):
    c = ex.candidates(l); r = ref["cands"][l]
    rc = np.stack([r["x"], r["y"], r["score"]], 1)
    cs = set(map(tuple, c.tolist())); rs = set(map(tuple, rc.tolist()))
    miss = sorted(rs - cs); extra = sorted(cs - rs)
    print("level", l, "dev", len(cs), "ref", len(rs), "missing", len(miss), "extra", len(extra))
    if l == 0:
        print(" missing sample", miss[:12]); print(" extra sample", extra[:12])
        m = np.array(miss)
        if len(m):
            print(" missing cell cols hist", np.bincount((m[:,0]-3)//36, minlength=35))
            print(" missing cell rows hist", np.bincount((m[:,1]-3)//37, minlength=19))
            print(" missing score hist", np.histogram(m[:,2], bins=[0,8,21,50,100,255])[0])
            cc = np.array(sorted(cs)); print(" dev score hist", np.histogram(cc[:,2], bins=[0,8,21,50,100,255])[0])
