"""CPU interpreter of the compiled "panel" refactor program (csparse3_b200/csrc/panel_program.hpp), test infrastructure.

Executes the step stream exactly as lu_panel.cu does (same accumulators, same multipliers, the L operands of an UPD
step read one step EARLY as the kernel's prefetch does) on a handful of systems at once, so the host compiler
(panel_program.cpp) is validated bit for bit against the oracle without a GPU.
"""
import numpy as np

from wide_interp import get_program

END, SCATTER, LOADU, UPD, PIV, FINU, FINL, NOP = range(8)
VALID, WS2, M0, M1, FUSED = 1, 2, 4, 8, 16


def decode(w):
    w = int(w)
    return dict(op=w >> 60, fl=(w >> 53) & 0x7f, c=(w >> 40) & 0x1fff, b=(w >> 20) & 0xfffff, a=w & 0xfffff,
                ab=w & 0xffffffffff)


def run_refactor(sym, Ax, fma=False):
    """Ax: [B, nnz] -> (Lx [B, lnz], Ux [B, unz], fail [B], stats)"""
    prog, geo = get_program(sym, 6)
    assert prog is not None, "panel program not available"
    _, width, NS, npanels, ops_decl, nsteps, smem, G = geo[:8]
    words = np.frombuffer(prog, dtype=np.uint64).reshape(-1, G)
    assert words.shape[0] == nsteps and smem == 2 * NS * width * 8
    B = Ax.shape[0]
    acc = np.zeros((2 * NS, B))
    Lg = np.full((sym.lnz, B), np.nan)
    Ug = np.full((sym.unz, B), np.nan)
    AxT = np.ascontiguousarray(Ax.T)
    fail = np.zeros(B, dtype=np.int64)
    u = np.zeros((2, 2, B))                 # u[source][acc]
    piv = np.ones(B); uk1 = np.zeros(B)
    ops = 0
    stats = dict(steps=0, ops=0, upd_rows=0)

    def fnma(x, l, m):
        return x - l * m        # numpy never fuses: bit-identical to __dsub_rn(x, __dmul_rn(l, m))

    pre = {}                                # L operands loaded one step ahead: (step, g) -> (l0, l1)
    for i in range(nsteps):
        recs = [decode(w) for w in words[i]]
        op = recs[0]["op"]
        assert all(r["op"] == op for r in recs), "mixed opcodes in step %d" % i
        # prefetch for step i + 1 happens before step i executes
        if i + 1 < nsteps:
            for g, w in enumerate(words[i + 1]):
                r = decode(w)
                if r["op"] == UPD and r["fl"] & VALID:
                    l0 = Lg[r["a"]].copy()
                    l1 = Lg[r["b"]].copy() if r["fl"] & WS2 else None
                    pre[(i + 1, g)] = (l0, l1)
        if op == END:
            break
        stats["steps"] += 1
        if op == SCATTER:
            dsts = [r["c"] for r in recs if r["fl"] & VALID]
            assert len(set(dsts)) == len(dsts)
            for r in recs:
                if r["fl"] & VALID:
                    acc[r["c"]] = AxT[r["ab"]]
        elif op == LOADU:
            r = recs[0]
            assert all(q == r for q in recs)
            for x, m in ((0, M0), (1, M1)):
                if r["fl"] & m:
                    u[0, x] = acc[x * NS + r["c"]]
                    if r["fl"] & WS2:
                        l = Lg[r["a"]]
                        assert not np.isnan(l).any()
                        u[1, x] = fnma(acc[x * NS + r["b"]], l, u[0, x])
                        acc[x * NS + r["b"]] = u[1, x]
                        ops += 1
        elif op == UPD:
            tg = [r["c"] for r in recs if r["fl"] & VALID]
            assert len(set(tg)) == len(tg)
            for g, r in enumerate(recs):
                if not r["fl"] & VALID:
                    continue
                l0, l1 = pre.pop((i, g))
                assert not np.isnan(l0).any(), "UPD reads an L entry that is not final (step %d)" % i
                stats["upd_rows"] += 1
                for x, m in ((0, M0), (1, M1)):
                    if r["fl"] & m:
                        t = x * NS + r["c"]
                        v = fnma(acc[t], l0, u[0, x]); ops += 1
                        if r["fl"] & WS2:
                            assert not np.isnan(l1).any()
                            v = fnma(v, l1, u[1, x]); ops += 1
                        acc[t] = v
        elif op == PIV:
            r = recs[0]
            piv = acc[r["c"]].copy()
            bad = ~((np.abs(piv) > 0) & np.isfinite(piv))
            fail[bad & (fail == 0)] = r["ab"]
            if r["fl"] & FUSED:
                assert r["c"] < NS
                uk1 = acc[NS + r["c"]].copy()
        elif op == FINU:
            for r in recs:
                if r["fl"] & VALID:
                    Ug[r["ab"]] = acc[r["c"]]
                    acc[r["c"]] = 0.0
        elif op == FINL:
            for r in recs:
                if r["fl"] & VALID:
                    with np.errstate(all="ignore"):
                        qv = acc[r["c"]] / piv
                    acc[r["c"]] = 0.0
                    Lg[r["ab"]] = qv
                    if r["fl"] & FUSED:
                        assert r["c"] < NS
                        acc[NS + r["c"]] = fnma(acc[NS + r["c"]], qv, uk1); ops += 1
        else:
            assert op == NOP
    assert not pre, "prefetched L operands never consumed"
    assert (acc == 0).all() or np.isnan(acc).any() or not np.isfinite(acc).all(), "accumulator not cleared at the end"
    Lg[sym.Lp[:-1]] = 1.0
    stats["ops"] = ops
    assert ops == ops_decl
    return np.ascontiguousarray(Lg.T), np.ascontiguousarray(Ug.T), fail, stats
