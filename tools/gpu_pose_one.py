"""One gpba_pose_optimize call on a seeded batch (for ncu)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
from pygpba import pose as PO
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 148
B = PO.make_pose_batch(n_frames=n_frames, n_pt=12000, A=2, outliers=0.1, seed=91, fix_prev=True)
R = PO.pose_optimize(B)
print("frames", B.n_frames, "matches", B.n_obs, "inliers", int(R.n_inliers.sum()))
