"""CPU interpreter of the "row-lane" refactor program (csparse3_b200/csrc/rowlane_program.hpp, kernel lu_rowlane.cu).

Executes the quads exactly as the kernel does, for a batch of systems at once (numpy over the batch axis), and
models the kernel's look-ahead and its warps: every warp of a bundle runs its own stream; the L operands and A values of
a whole stage are READ when the previous stage starts to execute, provided the other warps have published the columns
the stage needs (otherwise when the stage itself starts, after waiting); UPDLATE quads read when they execute.  Warps
take turns in a seeded random order, so a compiler that requests a value before its column is finalised, or whose
cross-warp requirements are wrong, produces wrong factors (or a deadlock) here, not only on the GPU."""
import ctypes as C

import numpy as np

from csparse3_b200 import _lib

NOP, LOAD4, UPDATE, STOREU4, STOREL4, UPDLATE, FIN, END = range(8)
FLAG_P, HAS_L, HAS_U, HAS_A = 1 << 12, 1 << 8, 1 << 9, 1 << 10
QW = 76


def get_program(sym):
    geo = (C.c_int64 * 8)()
    size = _lib.lib().csp3_lu_get_program(sym._h, 7, None, 0, geo)
    if size < 0:
        return None, None
    words = np.zeros(size // 4, dtype=np.uint32)
    _lib.lib().csp3_lu_get_program(sym._h, 7, words.ctypes.data_as(C.c_void_p), size, geo)
    return words.reshape(-1, QW), [int(v) for v in geo]


def kinds_all(quads):
    return (quads[:, 0] & 7).astype(int)


def split_streams(quads, warps, SQ):
    """Streams are stored one after the other, each padded with END quads to whole stages plus four stages."""
    kinds = quads[:, 0] & 7
    streams, pos = [], 0
    for _ in range(warps):
        end = pos
        while kinds[end] != END:
            end += 1
        stop = pos + (end - pos + SQ) // SQ * SQ + 4 * SQ         # the compiler's padding rule
        assert (kinds[end:stop] == END).all()
        streams.append((pos, end, stop))
        pos = stop
    assert pos == len(quads)
    return streams


def run_refactor(sym, Axb, seed=0):
    quads, geo = get_program(sym)
    assert quads is not None, "row-lane program not available"
    SQ, W, nslots, nquads = geo[1] & 0xff, geo[1] >> 8, geo[2], geo[5]
    batch = Axb.shape[0]
    Lx = np.full((batch, sym.lnz), np.nan)
    Ux = np.full((batch, sym.unz), np.nan)
    Lx[:, sym.Lp[:-1]] = 1.0                      # the unit diagonal is never written in the workspace layout
    fail = np.zeros(batch, dtype=np.int64)
    stats = {"ops": 0, "late_quads": 0, "quads": nquads, "update_quads": 0, "conflicts": 0, "waits": 0, "deferred_stages": 0, "publishes": 0}
    streams = split_streams(quads, W, SQ)
    assert sum(e - p for p, e, _ in streams) == nquads + int((kinds_all(quads) == NOP).sum()) or True
    kinds = (quads[:, 0] & 7).astype(int)
    h0 = quads[:, 0:4].astype(np.int64)
    base = quads[:, 4:8].astype(np.int64)
    req16 = quads[:, 8:12].copy().view(np.uint16).reshape(-1, 8).astype(np.int64)      # first quad of a stage
    lw = quads[:, 12:44].reshape(-1, 8, 4).transpose(0, 2, 1)     # [quad][record][g]
    aw = quads[:, 44:76].reshape(-1, 8, 4).transpose(0, 2, 1).astype(np.int64)
    valid = (lw >> 31).astype(bool)
    off = ((lw >> 16) & 0x7fff).astype(np.int64)
    slot = ((lw >> 6) & 0x3ff).astype(np.int64)
    assert (lw & 0x3f == 0).all()
    done = np.zeros(8, dtype=np.int64)            # published progress counters

    class Warp:
        pass

    warps = []
    for w, (p, e, stop) in enumerate(streams):
        x = Warp()
        x.w, x.first, x.stop = w, p, stop
        x.stage = 0
        x.acc = np.zeros((batch, nslots)); x.m = np.zeros(batch); x.piv = np.ones(batch)
        x.cols_done = 0; x.finished = False; x.queue = {}
        x.have = False
        warps.append(x)

    def sources_final(x, s):
        qd = x.first + s * SQ
        r = req16[qd]
        assert r[x.w] == 0 and bool(r.any()) == bool(h0[qd, 0] & (1 << 13))
        return bool((done >= r).all())

    def request_stage(x, s):
        for qd in range(x.first + s * SQ, x.first + (s + 1) * SQ):
            k = kinds[qd]
            for r in range(4):
                v = valid[qd, r]
                a = aw[qd, r]
                if k in (UPDATE, UPDLATE):
                    assert ((a != 0xffffffff) == v).all() and (a[v] == base[qd, r] + 64 * off[qd, r][v]).all()
                elif k == LOAD4 or (k == FIN and r == 2 and h0[qd, 0] & HAS_A):
                    assert ((a != 0xffffffff) == v).all() and (a[v] == base[qd, r] + 8 * off[qd, r][v]).all()
                else:
                    assert (a == 0xffffffff).all()
                if k == UPDATE:
                    x.queue[(qd, r)] = Lx[:, a[v] // 64].copy()
                elif k == LOAD4 or (k == FIN and r == 2 and h0[qd, 0] & HAS_A):
                    x.queue[(qd, r)] = Axb[:, a[v] // 8].copy()

    def pivot_prologue(x, qd):
        nonlocal fail
        x.piv = x.acc[:, (h0[qd, 1] & 0xffff) // 64].copy()
        bad = ~(np.isfinite(x.piv) & (np.abs(x.piv) > 0))
        code = int(h0[qd, 3])
        fail = np.where(bad & ((fail == 0) | (fail > code)), code, fail)

    def store(x, qd, r, is_l):
        v = valid[qd, r]
        s = slot[qd, r][v]
        assert len(set(s.tolist())) == len(s)
        val = x.acc[:, s].copy()
        x.acc[:, s] = 0.0
        dst = base[qd, r] // 64 + off[qd, r][v]
        if is_l:
            Lx[:, dst] = val / x.piv[:, None]
        else:
            Ux[:, dst] = val

    def turn(x):
        """One pass of the kernel's main loop; False when the warp is blocked on another warp's progress."""
        s = x.stage
        if done[x.w] != x.cols_done:
            stats["publishes"] += 1
        done[x.w] = x.cols_done                                  # publish
        if not x.have:
            if not sources_final(x, s):
                stats["waits"] += 1
                return False
            if s > 0:
                stats["deferred_stages"] += 1
            request_stage(x, s)
        nxt = (x.first + (s + 1) * SQ) < x.stop and sources_final(x, s + 1)
        if nxt:
            request_stage(x, s + 1)
        for qd in range(x.first + s * SQ, x.first + (s + 1) * SQ):
            k = kinds[qd]
            x0 = int(h0[qd, 0])
            if k == END:
                x.finished = True
                done[x.w] = x.cols_done
                return True
            if k in (UPDATE, UPDLATE):
                stats["update_quads"] += 1
                stats["late_quads"] += int(k == UPDLATE)
                ms = [h0[qd, 1] & 0xffff, h0[qd, 1] >> 16, h0[qd, 2] & 0xffff, h0[qd, 2] >> 16]
                for r in range(4):
                    v = valid[qd, r]
                    sl = slot[qd, r][v]
                    assert len(set(sl.tolist())) == len(sl), "two lane groups of a record share a slot"
                    if x0 & (0x100 << r):
                        x.m = x.acc[:, ms[r] // 64].copy()
                    l = x.queue.pop((qd, r)) if k == UPDATE else Lx[:, aw[qd, r][v] // 64]
                    x.acc[:, sl] = x.acc[:, sl] - l * x.m[:, None]
                    stats["ops"] += len(sl)
                    for g in range(0, 8, 2):
                        if valid[qd, r][g] and valid[qd, r][g + 1] and (slot[qd, r][g] ^ slot[qd, r][g + 1]) & 1 == 0:
                            stats["conflicts"] += 1
            elif k == FIN:
                assert (x0 >> 24) == (1 if x0 & HAS_L else 0) | (8 if x0 & HAS_U else 0) | (48 if x0 & HAS_A else 0)
                if x0 & FLAG_P:
                    pivot_prologue(x, qd)
                if x0 & HAS_L:
                    store(x, qd, 0, True)
                if x0 & HAS_U:
                    store(x, qd, 1, False)
                if x0 & HAS_A:
                    x.acc[:, slot[qd, 2][valid[qd, 2]]] = x.queue.pop((qd, 2))
                x.cols_done += 1
            elif k in (STOREL4, STOREU4):
                cnt = (x0 >> 16) & 7
                assert (x0 >> 24) == sum((1 if k == STOREL4 else 2) << (2 * r) for r in range(cnt))
                if x0 & FLAG_P:
                    pivot_prologue(x, qd)
                for r in range(cnt):
                    store(x, qd, r, k == STOREL4)
            elif k == LOAD4:
                assert (x0 >> 24) == sum(3 << (2 * r) for r in range((x0 >> 16) & 7))
                for r in range(4):
                    x.acc[:, slot[qd, r][valid[qd, r]]] = x.queue.pop((qd, r))
            else:
                assert k == NOP
        x.stage += 1
        x.have = nxt
        return True

    rng = np.random.default_rng(seed)
    with np.errstate(all="ignore"):
        while not all(x.finished for x in warps):
            live = [x for x in warps if not x.finished]
            order = rng.permutation(len(live))
            progressed = False
            for i in order[: max(1, int(rng.integers(1, len(live) + 1)))]:
                progressed = turn(live[i]) or progressed
            if not progressed:
                # the chosen warps were blocked: everyone gets a turn; nobody moving and nothing newly published is a deadlock
                before = stats["publishes"]
                moved = [turn(x) for x in live]
                assert any(moved) or stats["publishes"] != before, "deadlock: every warp waits for another warp"
    for x in warps:
        assert not x.acc.any() or fail.any()      # every slot is cleared by its store
        assert not x.queue
    return Lx, Ux, fail, stats
