"""Phase clocks of k_chol_panel (needs tools/_dbg/libgpba_timing.so built with -DGPBA_CHOL_TIMING)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "amc-slam_b200"))
from pygpba import synth, lib as G
G.LIB_PATH = os.path.join(ROOT, "tools", "_dbg", "libgpba_timing.so")
P = synth.make_problem("c4")
g = G.GpBa(P)
g.optimize(2)
clk = (C.c_longlong * 64)()
G.lib().gpba_debug_chol_clocks(clk)
v = list(clk)
print("raw", v[:10])
prev = v[1]
for pb in range(6):
    t = v[16 + 4 * pb: 20 + 4 * pb]
    print(f"potrf pb={pb}: load+bar {t[0]-prev} factor {t[1]-t[0]} write+sync {t[2]-t[1]} trailing+sync {t[3]-t[2]}")
    prev = t[3]
t0 = v[16 + 20]
print("pb=5 thread 0  :", [x - t0 for x in v[36:40]])
print("pb=5 thread 74 :", [x - t0 for x in v[40:44]])
print("pb=5 thread 160:", [x - t0 for x in v[44:48]])
print("pb=5 thread 64 :", [x - t0 for x in v[48:52]])
print("phases (cycles):", [v[i + 1] - v[i] for i in range(4)], "total", v[4] - v[0])
