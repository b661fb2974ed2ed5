"""CPU interpreter of the compiled "panel" refactor program (csparse3_b200/csrc/panel_program.hpp), test infrastructure.

Executes the step stream exactly as lu_panel.cu does -- same accumulators, same ring / landing area, same fetch
timing -- on a handful of systems at once (numpy vectors over the batch), so the host compiler (panel_program.cpp)
is validated bit for bit against the oracle without a GPU.  cp.async is modelled at its two extremes: the destination
is poisoned (NaN) when a fetch is issued, because the data may land at any moment from then on, and the data only
becomes readable LOOKAHEAD steps later.
"""
import numpy as np

from wide_interp import get_program

NONE, SCATTER, HDRU, UPD, HDRP, FINU, FINL, FETCHL, FETCHA, END = 0, 1, 2, 3, 4, 5, 6, 7, 8, 15
WS2, M0, M1, FUSED, X, Y = 2, 4, 8, 16, 32, 64
LOOKAHEAD = 8
STEP_WORDS = 16


def decode(w):
    w = int(w)
    return dict(op=w >> 60, fl=(w >> 53) & 0x7f, c=(w >> 40) & 0x1fff, b=(w >> 20) & 0xfffff, a=w & 0xfffff,
                ab=w & 0xffffffffff)


def run_refactor(sym, Ax):
    """Ax: [B, nnz] -> (Lx [B, lnz], Ux [B, unz], fail [B], stats)"""
    prog, geo = get_program(sym, 6)
    assert prog is not None, "panel program not available"
    R, width, NS, LN, ops_decl, nsteps, smem, G = geo[:8]
    words = np.frombuffer(prog, dtype=np.uint64).reshape(-1, STEP_WORDS)
    assert words.shape[0] == nsteps and nsteps % 4 == 0
    assert smem == (2 * NS + R + LN) * width * 8 + 4 * 4 * STEP_WORDS * 8
    B = Ax.shape[0]
    acc = np.zeros((2 * NS, B))
    lsrc = np.full((R + LN, B), np.nan)
    Lg = np.full((sym.lnz, B), np.nan)
    Ug = np.full((sym.unz, B), np.nan)
    AxT = np.ascontiguousarray(Ax.T)
    fail = np.zeros(B, dtype=np.int64)
    u = np.zeros((2, 2, B))                 # u[source][acc]
    piv = np.ones(B); uk1 = np.zeros(B)
    tfl = 0
    ops = 0
    pending = []                            # (ready step, lsrc entry, data)
    stats = dict(steps=0, ops=0, upd_rows=0, fetches=0, sync=0)
    ended = False

    def fnma(x, l, m):
        return x - l * m        # numpy never fuses: bit-identical to __dsub_rn(x, __dmul_rn(l, m))

    def lval(idx, sync):
        if sync:
            v = Lg[idx]
            stats["sync"] += 1
        else:
            v = lsrc[idx]
        assert not np.isnan(v).any(), "L operand not available (step %d)" % i
        return v

    for i in range(nsteps):
        recs = [decode(w) for w in words[i]]
        # fetches first: poison now, data readable LOOKAHEAD steps later
        for r in recs:
            if r["op"] in (FETCHL, FETCHA):
                e = r["c"]
                assert R <= e < R + LN, "fetch outside the landing area"
                if r["op"] == FETCHL:
                    data = Lg[r["ab"]].copy()
                    assert not np.isnan(data).any(), "fetch of an L entry that is not final (step %d)" % i
                else:
                    data = AxT[r["ab"]].copy()
                lsrc[e] = np.nan
                pending.append((i + LOOKAHEAD, e, data))
                stats["fetches"] += 1
        keep = []
        for ready, e, data in pending:
            if ready <= i:
                lsrc[e] = data
            else:
                keep.append((ready, e, data))
        pending = keep
        hd = recs[0]
        if hd["op"] == END:
            ended = True
            break
        stats["steps"] += 1
        if hd["op"] == HDRU:
            tfl = hd["fl"]
            for x, m in ((0, M0), (1, M1)):
                if tfl & m:
                    u[0, x] = acc[x * NS + hd["c"]]
                    if tfl & WS2:
                        l = lval(hd["a"], tfl & X)
                        u[1, x] = fnma(acc[x * NS + hd["b"]], l, u[0, x])
                        acc[x * NS + hd["b"]] = u[1, x]
                        ops += 1
        elif hd["op"] == HDRP:
            piv = acc[hd["c"]].copy()
            bad = ~((np.abs(piv) > 0) & np.isfinite(piv))
            fail[bad & (fail == 0)] = hd["ab"]
            if hd["fl"] & FUSED:
                assert hd["c"] < NS
                uk1 = acc[NS + hd["c"]].copy()
        main = set(r["op"] for r in recs[(1 if hd["op"] in (HDRU, HDRP) else 0):]) - {NONE, FETCHL, FETCHA}
        assert len(main) <= 1 or main == {FINU, FINL}, "mixed step %d: %s" % (i, main)
        upd = [r for r in recs if r["op"] == UPD]
        if upd:
            tg = [r["c"] for r in upd]
            assert len(set(tg)) == len(tg)
            # all loads, then all stores (the kernel's lanes run in lockstep)
            vals = []
            for r in upd:
                l0 = lval(r["a"], r["fl"] & X).copy()
                l1 = lval(r["b"], r["fl"] & Y).copy() if tfl & WS2 else None
                vals.append((l0, l1, [acc[x * NS + r["c"]].copy() for x in (0, 1)]))
            for r, (l0, l1, xs) in zip(upd, vals):
                stats["upd_rows"] += 1
                for x, m in ((0, M0), (1, M1)):
                    if tfl & m:
                        v = fnma(xs[x], l0, u[0, x]); ops += 1
                        if tfl & WS2:
                            v = fnma(v, l1, u[1, x]); ops += 1
                        acc[x * NS + r["c"]] = v
        for r in recs:
            if r["op"] == SCATTER:
                if r["fl"] & X:
                    acc[r["c"]] = AxT[r["ab"]]; stats["sync"] += 1
                else:
                    v = lsrc[r["ab"]]
                    assert not np.isnan(v).any() or np.isnan(AxT).any(), "A value has not landed (step %d)" % i
                    acc[r["c"]] = v
            elif r["op"] == FINU:
                Ug[r["ab"]] = acc[r["c"]]
                acc[r["c"]] = 0.0
            elif r["op"] == FINL:
                with np.errstate(all="ignore"):
                    qv = acc[r["c"]] / piv
                acc[r["c"]] = 0.0
                Lg[r["ab"] & 0xffffff] = qv
                if r["fl"] & X:
                    assert (r["ab"] >> 24) < R
                    lsrc[r["ab"] >> 24] = qv
                if r["fl"] & FUSED:
                    assert r["c"] < NS
                    acc[NS + r["c"]] = fnma(acc[NS + r["c"]], qv, uk1); ops += 1
    assert ended
    Lg[sym.Lp[:-1]] = 1.0
    stats["ops"] = ops
    assert ops == ops_decl
    return np.ascontiguousarray(Lg.T), np.ascontiguousarray(Ug.T), fail, stats
