"""Static SASS instruction counts per kernel of amc-slam_b200/libgpba.so (cuobjdump -sass): which kernels carry FP64 tensor-core MMAs
(DMMA), TMA bulk copies (UBLKCP) and mbarrier operations (SYNCS).  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "amc-slam_b200", "libgpba.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
PAT = (("DMMA", r"\bDMMA\b"), ("DFMA", r"\bDFMA\b"), ("UBLKCP", r"\bUBLKCP\b"), ("SYNCS", r"\bSYNCS\b"), ("MUFU", r"\bMUFU\b"), ("LDG", r"\bLDG\b"),
       ("LDS", r"\bLDS\b"), ("STG", r"\bSTG\b"), ("ATOM/RED", r"\bATOMG?\b|\bRED\b|\bREDG\b"), ("MEMBAR", r"\bMEMBAR\b"))
cur = None
cnt = collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    if cur:
        for key, pat in PAT:
            if re.search(pat, line):
                cnt[cur][key] += 1
print("SASS summary of amc-slam_b200/libgpba.so (cuobjdump -sass, sm_100a; static instruction counts per kernel).")
print("DMMA = FP64 tensor-core MMA (every f64 mma.sync shape lowers to DMMA.8x8x4 on sm_100a); UBLKCP = cp.async.bulk (TMA bulk copy);")
print("SYNCS = mbarrier operations (ARRIVE.TRANS64 = expect_tx, PHASECHK = try_wait); MEMBAR = gpu-scope fences of the dataflow kernel.\n")
print(f"{'kernel':40s}" + "".join(f"{k:>9s}" for k, _ in PAT))
tot = collections.Counter()
for k in sorted(cnt, key=lambda k: subprocess.run(["c++filt", k], capture_output=True, text=True).stdout):
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    if "gpba::" not in name and "k_" not in name:
        continue
    short = re.sub(r"\(.*", "", name).replace("gpba::", "").replace("void ", "").replace("(anonymous namespace)::", "")
    c = cnt[k]
    print(f"{short[:40]:40s}" + "".join(f"{c[key]:9d}" for key, _ in PAT))
    tot.update(c)
print(f"{'TOTAL':40s}" + "".join(f"{tot[key]:9d}" for key, _ in PAT))
