"""CPU interpreter of the "row-lane" refactor program (csparse3_b200/csrc/rowlane_program.hpp, kernel lu_rowlane.cu).

Executes the quads exactly as the kernel does, for a batch of systems at once (numpy over the batch axis), and
models the kernel's look-ahead: the L operands and A values of a whole stage (4 quads) are READ when the previous stage
starts to execute (UPDLATE quads read when they execute), so a compiler that requests a value before its column is
finalised produces wrong factors here, not only on the GPU."""
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


def run_refactor(sym, Axb):
    quads, geo = get_program(sym)
    assert quads is not None, "row-lane program not available"
    SQ, nslots, nquads = geo[1], geo[2], geo[5]
    batch = Axb.shape[0]
    Lx = np.full((batch, sym.lnz), np.nan)
    Ux = np.full((batch, sym.unz), np.nan)
    Lx[:, sym.Lp[:-1]] = 1.0                      # the unit diagonal is never written in the workspace layout
    acc = np.zeros((batch, nslots))
    fail = np.zeros(batch, dtype=np.int64)
    m = np.zeros(batch)
    piv = np.ones(batch)
    stats = {"ops": 0, "late_quads": 0, "quads": nquads, "update_quads": 0, "conflicts": 0}
    assert len(quads) % SQ == 0 and len(quads) >= nquads + 4 * SQ and (quads[nquads:, 0] & 7 == END).all()
    kinds = (quads[:, 0] & 7).astype(int)
    h0 = quads[:, 0:4].astype(np.int64)
    base = quads[:, 4:8].astype(np.int64)
    lw = quads[:, 12:44].reshape(-1, 8, 4).transpose(0, 2, 1)     # [quad][record][g]
    aw = quads[:, 44:76].reshape(-1, 8, 4).transpose(0, 2, 1).astype(np.int64)
    valid = (lw >> 31).astype(bool)
    off = ((lw >> 16) & 0x7fff).astype(np.int64)
    slot = ((lw >> 6) & 0x3ff).astype(np.int64)
    assert (lw & 0x3f == 0).all()
    queue = {}                                     # (quad, record) -> operand values requested one stage ahead

    def request_stage(s):
        for qd in range(s * SQ, min((s + 1) * SQ, len(quads))):
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
                    queue[(qd, r)] = Lx[:, a[v] // 64].copy()
                elif k == LOAD4 or (k == FIN and r == 2 and h0[qd, 0] & HAS_A):
                    queue[(qd, r)] = Axb[:, a[v] // 8].copy()

    def pivot_prologue(qd):
        nonlocal piv, fail
        piv = acc[:, (h0[qd, 1] & 0xffff) // 64].copy()
        bad = ~(np.isfinite(piv) & (np.abs(piv) > 0))
        code = int(h0[qd, 3])
        fail = np.where(bad & ((fail == 0) | (fail > code)), code, fail)

    def store(qd, r, is_l):
        v = valid[qd, r]
        s = slot[qd, r][v]
        assert len(set(s.tolist())) == len(s)
        x = acc[:, s].copy()
        acc[:, s] = 0.0
        dst = base[qd, r] // 64 + off[qd, r][v]
        if is_l:
            Lx[:, dst] = x / piv[:, None]
        else:
            Ux[:, dst] = x

    request_stage(0)
    done = False
    with np.errstate(all="ignore"):
        for s in range(len(quads) // SQ):
            request_stage(s + 1)
            for qd in range(s * SQ, (s + 1) * SQ):
                k = kinds[qd]
                x0 = int(h0[qd, 0])
                if k == END:
                    assert qd == nquads
                    done = True
                    break
                if k in (UPDATE, UPDLATE):
                    stats["update_quads"] += 1
                    stats["late_quads"] += int(k == UPDLATE)
                    ms = [h0[qd, 1] & 0xffff, h0[qd, 1] >> 16, h0[qd, 2] & 0xffff, h0[qd, 2] >> 16]
                    for r in range(4):
                        v = valid[qd, r]
                        sl = slot[qd, r][v]
                        assert len(set(sl.tolist())) == len(sl), "two lane groups of a record share a slot"
                        if x0 & (0x100 << r):
                            m = acc[:, ms[r] // 64].copy()
                        l = queue.pop((qd, r)) if k == UPDATE else Lx[:, aw[qd, r][v] // 64]
                        acc[:, sl] = acc[:, sl] - l * m[:, None]
                        stats["ops"] += len(sl)
                        for g in range(0, 8, 2):
                            if valid[qd, r][g] and valid[qd, r][g + 1] and (slot[qd, r][g] ^ slot[qd, r][g + 1]) & 1 == 0:
                                stats["conflicts"] += 1
                elif k == FIN:
                    assert (x0 >> 24) == (1 if x0 & HAS_L else 0) | (8 if x0 & HAS_U else 0) | (48 if x0 & HAS_A else 0)
                    if x0 & FLAG_P:
                        pivot_prologue(qd)
                    if x0 & HAS_L:
                        store(qd, 0, True)
                    if x0 & HAS_U:
                        store(qd, 1, False)
                    if x0 & HAS_A:
                        acc[:, slot[qd, 2][valid[qd, 2]]] = queue.pop((qd, 2))
                elif k in (STOREL4, STOREU4):
                    cnt = (x0 >> 16) & 7
                    assert (x0 >> 24) == sum((1 if k == STOREL4 else 2) << (2 * r) for r in range(cnt))
                    if x0 & FLAG_P:
                        pivot_prologue(qd)
                    for r in range((x0 >> 16) & 7):
                        store(qd, r, k == STOREL4)
                elif k == LOAD4:
                    assert (x0 >> 24) == sum(3 << (2 * r) for r in range((x0 >> 16) & 7))
                    for r in range(4):
                        acc[:, slot[qd, r][valid[qd, r]]] = queue.pop((qd, r))
                else:
                    assert k == NOP
            if done:
                break
    assert done
    assert not acc.any() or fail.any()            # every slot is cleared by its store
    return Lx, Ux, fail, stats
