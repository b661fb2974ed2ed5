"""Structure of a cached LU pattern: where the refactor arithmetic is (per column, per supernode, trailing block)."""
import sys
import numpy as np
from csparse3_b200 import synth
from csparse3_b200.lu import LuSymbolic

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
g = synth.GridCase(nb)
n, Ap, Ai, Ax0 = g.base_jacobian()
sym = LuSymbolic(n, Ap, Ai, Ax0)
Lp, Li, Up, Ui = sym.Lp, sym.Li, sym.Up, sym.Ui
lcnt = np.diff(Lp) - 1          # below-diagonal entries per L column
ucnt = np.diff(Up) - 1          # above-diagonal entries per U column
ops = np.zeros(n, dtype=np.int64)
for k in range(n):
    js = Ui[Up[k]:Up[k + 1] - 1]
    ops[k] = lcnt[js].sum()
print("n", n, "nnzA", sym.nnz, "lnz", sym.lnz, "unz", sym.unz, "ops", ops.sum(), "flops", sym.flops)
order = np.argsort(-ops)
cs = np.cumsum(ops[order])
for frac in (0.5, 0.8, 0.9, 0.95, 0.99):
    print("columns holding %.0f%% of ops: %d" % (frac * 100, np.searchsorted(cs, frac * cs[-1]) + 1))
# trailing block: ops in last t columns
for t in (32, 64, 96, 128, 192, 256, 384, 512, 1024):
    print("last %4d columns: ops %.3f of total, max ucnt %d max lcnt %d" % (t, ops[n - t:].sum() / ops.sum(), ucnt[n-t:].max(), lcnt[n-t:].max()))
# fundamental supernodes in L: column j+1 pattern == column j pattern minus row j+1 (rows in pivot order)
pinv = sym.pinv
rows = [np.sort(Li[Lp[j] + 1:Lp[j + 1]]) for j in range(n)]  # Li already in pivotal numbering? check
sn_start = [0]
for j in range(1, n):
    a, b = rows[j - 1], rows[j]
    if len(a) == len(b) + 1 and a[0] == j and np.array_equal(a[1:], b):
        continue
    sn_start.append(j)
sn_start.append(n)
w = np.diff(sn_start)
print("supernodes", len(w), "width hist", np.bincount(np.minimum(w, 20)))
# ops by source supernode width
opsw = np.zeros(64, dtype=np.int64)
snof = np.repeat(np.arange(len(w)), w)
for k in range(n):
    js = Ui[Up[k]:Up[k + 1] - 1]
    for j in js:
        opsw[min(w[snof[j]], 63)] += lcnt[j]
print("ops by source-supernode width:", {i: int(v) for i, v in enumerate(opsw) if v})
print("etree-ish: lcnt top 20 cols", lcnt[-20:], "ucnt", ucnt[-20:])
print("max lcnt", lcnt.max(), "max ucnt", ucnt.max())

# ---- reuse of L columns: LRU model per system (capacity in factor entries) -----------------------------
# access trace: for target column k: for j in U(:,k) (stored order): read L(:,j) [lcnt_j entries]; then write L(:,k),U(:,k)
from collections import OrderedDict
def lru_misses(cap):
    lru = OrderedDict(); used = 0; miss = 0; tot = 0
    for k in range(n):
        for j in Ui[Up[k]:Up[k + 1] - 1]:
            sz = int(lcnt[j])
            if sz == 0: continue
            tot += sz
            if j in lru:
                lru.move_to_end(j)
            else:
                miss += sz
                lru[j] = sz; used += sz
        sz = int(lcnt[k])
        if sz:
            lru[k] = sz; used += sz; lru.move_to_end(k)
        while used > cap:
            _, s = lru.popitem(last=False); used -= s
    return miss, tot
for cap in (256, 512, 1024, 2048, 4096, 8192, 16384):
    m, t = lru_misses(cap)
    print("LRU cap %5d entries (%4d KB/system): L re-read misses %6d entries of %d accesses (%.2fx of lnz)" % (cap, cap * 8 // 1024, m, t, m / sym.lnz))
