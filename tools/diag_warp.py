import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import align as oa
print(open('/proc/cpuinfo').read().split('model name')[1].split('\n')[0])
print(cv2.__version__, cv2.getNumThreads())
rng = np.random.default_rng(212)
def case(S, H=200, W=240):
    src = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), 2)
    tpl = oa.template(S)
    ang = np.deg2rad(rng.uniform(-25, 25)); sc = rng.uniform(0.6, 2.2) * 112 / S
    R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]) * sc
    lm = (tpl - S / 2) @ R.T + np.array([W / 2 + rng.uniform(-70, 70), H / 2 + rng.uniform(-70, 70)]) + rng.normal(0, 0.5, (5, 2))
    return src, lm
from facerecognitionpipeline_b200.face_recognition import FaceAligner
al = FaceAligner(112)
for t in range(5):
    src, lm = case(112)
    M = oa.estimate(lm, 112)
    ref = oa.align(src, lm, 112)
    emu = oa.warp_affine_fixed_point(src, M, 112)
    got = al.align(src, lm)
    for name, a, b in (("cv2 vs emu", ref, emu), ("gpu vs emu", got, emu), ("gpu vs cv2", got, ref)):
        d = np.abs(a.astype(int) - b.astype(int))
        bad = np.argwhere(d.max(axis=2) > 0)
        print(t, name, "nbad", len(bad), "max", d.max(), "first", bad[:5].tolist())
    print("M", M.tolist())
