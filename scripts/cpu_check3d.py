"""CPU-only check of the 3D host pipeline (drop-in GPisMap3 linked against the oracle-backed mock
C ABI, tests/mockbuild.py) against the unmodified reference on the bundled BigBIRD frames.
Needs /root/reference (data) — a development aid, not a test."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, cv2
from oracle import refpy
from gpismap_b200 import hostapi
import mockbuild
mock = mockbuild.build()
nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 6
base='/root/reference/data/3D/bigbird_detergent'
poses = np.loadtxt(base+'/pose/poses.txt').astype(np.float32)
FrameNums = list(range(93,360,3))+list(range(3,91,3)); CamIDs = [1,2,3,4,3,2]*30
fx = [570.9361, 572.3318, 568.9403 , 567.9881, 572.7638]; fy=[570.9376, 572.3316, 568.9419 , 567.9995, 572.7567]
cx=[306.8789, 309.9968, 308.4583, 310.5243, 310.4192]; cy=[238.8476, 230.6296, 225.8232, 223.9443, 214.8762]
R = None; G=None
count=0
xg,yg,zg = np.meshgrid(np.arange(-0.07,0.1301,0.01), np.arange(-0.1,0.1401,0.01), np.arange(0,0.2801,0.01))
X = np.stack([xg.ravel(order='F'), yg.ravel(order='F'), zg.ravel(order='F')],1).astype(np.float32)
for k in range(0, len(FrameNums), 3):
    frm = FrameNums[k]; cam = CamIDs[count]; row = poses[count]; count+=1
    D = cv2.imread(f'{base}/masked_depth/frame{frm}_cam{cam}.png', cv2.IMREAD_UNCHANGED).astype(np.float32)*np.float32(0.0001)
    t = row[[3,7,11]]; Rm = row[[0,1,2,4,5,6,8,9,10]]
    pose = np.concatenate([t,Rm]).astype(np.float32)
    c = (np.float32(fx[cam-1]), np.float32(fy[cam-1]), np.float32(cx[cam-1]), np.float32(cy[cam-1]), 640, 480)
    if R is None:
        R = refpy.RefMap3(cam=c); G = hostapi.GPisMap3(cam=c, libpath=mock)
    else:
        R.set_cam(*c); G.resetCam(*c)
    dz = np.ascontiguousarray(D.T).ravel()
    t0=time.time(); R.update(dz, pose); t1=time.time(); G.update(dz, pose); t2=time.time()
    a = R.all_samples(); b = G.all_samples()
    ca, na, tr = R.clusters(); cb, nb = G.leaves()
    print(count, frm, cam, 'valid', (dz>0).sum(), 'samples', a.shape[0], b.shape[0], 'identical', a.shape==b.shape and np.array_equal(a,b),
          'leaves', len(ca), len(cb), 'ref %.2fs mine(mock) %.2fs'%(t1-t0,t2-t1), flush=True)
    if count>=nframes: break
ra = R.test(X); rb = G.test(X)
print('test rows differing', (np.abs(ra-rb).max(1)>0).sum(), 'of', len(X), 'max abs', np.abs(ra-rb).max())
