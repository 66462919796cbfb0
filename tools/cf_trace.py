"""Task timeline of the persistent factorization kernel (k_chol_factor), from GPBA_CF_TRACE=<file> dumps.

    GPBA_CF_TRACE=/tmp/cf.bin python tools/gpu_one.py c4 1 ; python tools/cf_trace.py /tmp/cf.bin [n_ctas]

Per task: claimed (producer took it from the list), ready (its last dependency wait returned), start (consumers saw its
first item), done (counted).  Prints where the time of one factorization goes: per level and per task kind."""
import sys
import numpy as np

path = sys.argv[1]
n_ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 148
raw = open(path, "rb").read()
n = int(np.frombuffer(raw[:4], np.int32)[0])
tab = np.frombuffer(raw[4:4 + 16 * n], np.int32).reshape(n, 4)
tr8 = np.frombuffer(raw[4 + 16 * n:], np.int64).reshape(n, 8).astype(np.float64)
tr = tr8[:, :4].copy()
t0 = tr[:, 0].min()
tr = (tr - t0) / 1e3   # us
panel = tab[:, 2] < 0
# level index: a new level starts where a chunk follows a panel task (level 0 has panel tasks only)
lvl = np.zeros(n, int)
cur = 0
for i in range(1, n):
    if panel[i - 1] and not panel[i]:
        cur += 1
    lvl[i] = cur
span = tr[:, 3].max()
print(f"{n} tasks ({panel.sum()} panel, {(~panel).sum()} chunks, {int((tab[~panel, 3] - tab[~panel, 2]).sum())} products), {cur + 1} levels, span {span:.1f} us")
for name, m in (("panel", panel), ("chunk", ~panel)):
    w = tr[m, 1] - tr[m, 0]; ld = tr[m, 2] - tr[m, 1]; ex = tr[m, 3] - tr[m, 2]
    print(f"  {name}: wait mean {w.mean():6.2f} us (sum {w.sum() / 1e3:7.2f} ms)   ready->start mean {ld.mean():5.2f}   start->done mean {ex.mean():5.2f} (sum {ex.sum() / 1e3:6.2f} ms)")
ch = ~panel
nprod = (tab[ch, 3] - tab[ch, 2])
print(f"  chunk detail: products/chunk {nprod.mean():.2f}; consumer wait-for-operands inside chunk mean {tr8[ch, 4].mean() / 1e3:.2f} us; "
      f"compute end -> counted mean {(tr8[ch, 3] - tr8[ch, 5]).mean() / 1e3:.2f} us; producer wait-for-free-stage mean {tr8[ch, 6].mean() / 1e3:.2f} us")
pm = panel & (tr8[:, 4] > 0)   # tasks that factorize (solve-only tasks of wide levels leave the slot empty)
print(f"  panel detail: start -> potrf done mean {(tr8[pm, 4] - tr8[pm, 2]).mean() / 1e3:.2f} us, potrf -> trsm done {(tr8[pm, 5] - tr8[pm, 4]).mean() / 1e3:.2f} us, trsm -> counted {(tr8[pm, 3] - tr8[pm, 5]).mean() / 1e3:.2f} us")
busy = (tr[:, 3] - tr[:, 2]).sum()
print(f"  consumer-busy fraction over {n_ctas} CTAs: {busy / (span * n_ctas):.3f}")
print(" level  chunks panels   first_claim  chunks_done  panels_done   level_len")
prev_done = 0.0
rows = []
for l in range(cur + 1):
    m = lvl == l
    pc = m & panel; cc = m & ~panel
    cd = tr[cc, 3].max() if cc.any() else float("nan")
    pd = tr[pc, 3].max()
    rows.append((l, int(cc.sum()), int(pc.sum()), tr[m, 0].min(), cd, pd, pd - prev_done))
    prev_done = pd
for r in rows:
    print(f"{r[0]:6d} {r[1]:7d} {r[2]:6d} {r[3]:12.1f} {r[4]:12.1f} {r[5]:12.1f} {r[6]:10.1f}")
