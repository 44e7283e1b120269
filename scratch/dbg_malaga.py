import sys, os; sys.path.insert(0,'.')
import numpy as np
from monocular_visual_odometry_va4mr_b200 import synth, workload, cv2_compat
o=workload.REFERENCE_OPTIONS["malaga"]
s=synth.render_sequence("malaga",2,seed=4); f0,f1=s["frames"]
pts=synth.grid_corners(f0,1000,seed=2)
q=np.ascontiguousarray(pts[537:538])
p,st,err=cv2_compat.calcOpticalFlowPyrLK(f0,f1,q,None,winSize=o["win"],maxLevel=5,criteria=o["criteria"])
print(p, st, err)
