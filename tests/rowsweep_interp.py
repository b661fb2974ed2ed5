"""CPU interpreter of the row-sweep programs (csparse3_b200/csrc/rowsweep_program.hpp, kernel lu_sweep_rows_kernel):
level by level, every warp's panels of the level in a seeded random order (rows of a level are independent), per row the
terms in program order with unfused multiply / subtract.  Checks along the way that a term never reads a row of the same
or a later level."""
import ctypes as C

import numpy as np

from csparse3_b200 import _lib

NONE, ENDLEVEL, HDR, CHUNK = 0xFFFFFFFF, 0x7FFFFFFF, 20, 128


def get_program(sym, which):
    geo = (C.c_int64 * 8)()
    size = _lib.lib().csp3_lu_get_program(sym._h, which, None, 0, geo)
    if size < 0:
        return None, None
    words = np.zeros(size // 4, dtype=np.uint32)
    _lib.lib().csp3_lu_get_program(sym._h, which, words.ctypes.data_as(C.c_void_p), size, geo)
    return words, [int(v) for v in geo]


def panels_of(words, warps):
    """-> per warp: list of (level, rowoff[8], diagoff[8], terms[chunks * 8 steps][8 groups][2])"""
    out, pos = [], 0
    for _ in range(warps):
        lst = []
        while True:
            level, nch = int(words[pos]), int(words[pos + 1])
            if level == ENDLEVEL:
                pos += HDR
                break
            ro = words[pos + 4:pos + 12].astype(np.int64)
            dg = words[pos + 12:pos + 20].astype(np.int64)
            tw = words[pos + HDR:pos + HDR + nch * CHUNK].astype(np.int64).reshape(nch * 8, 8, 2)
            lst.append((level, ro, dg, tw))
            pos += HDR + nch * CHUNK
        out.append(lst)
    assert pos == len(words)
    return out


def sweep(sym, which, Fx, z, seed=0):
    """In place on z[batch, n] (pivot order); Fx[batch, entries] the factor values."""
    words, geo = get_program(sym, which)
    assert words is not None
    levels, warps = geo[0], geo[1]
    streams = panels_of(words, warps)
    lower = which == 9
    rng = np.random.default_rng(seed)
    level_of_row = np.full(z.shape[1], -1)
    for st in streams:
        for level, ro, dg, tw in st:
            for g in range(8):
                if ro[g] != NONE:
                    assert ro[g] % 64 == 0 and level_of_row[ro[g] // 64] < 0
                    level_of_row[ro[g] // 64] = level
    assert (level_of_row >= 0).all()
    cursor = [0] * warps
    terms = 0
    with np.errstate(all="ignore"):
        for lv in range(levels):
            todo = []
            for w in range(warps):
                while cursor[w] < len(streams[w]) and streams[w][cursor[w]][0] == lv:
                    todo.append(streams[w][cursor[w]]); cursor[w] += 1
            new = {}
            for k in rng.permutation(len(todo)):
                level, ro, dg, tw = todo[k]
                for g in range(8):
                    if ro[g] == NONE:
                        assert (tw[:, g, 0] == NONE).all()
                        continue
                    i = ro[g] // 64
                    acc = z[:, i].copy()
                    seen_none = False
                    for t in range(tw.shape[0]):
                        fo, yo = tw[t, g]
                        if fo == NONE:
                            seen_none = True
                            continue
                        assert not seen_none and fo % 64 == 0 and yo % 64 == 0
                        j = yo // 64
                        assert level_of_row[j] < lv, "a term reads a row that is not final"
                        acc = acc - Fx[:, fo // 64] * z[:, j]
                        terms += 1
                    if not lower:
                        assert dg[g] % 64 == 0
                        acc = acc / Fx[:, dg[g] // 64]
                    new[i] = acc
            for i, v in new.items():                   # results of a level become visible at the barrier
                z[:, i] = v
    assert all(cursor[w] == len(streams[w]) for w in range(warps))
    assert terms == geo[4]
    return z


def solve(sym, Lx, Ux, b, seed=0):
    """x = A \\ b through the two row-sweep programs (batch axis first)."""
    n = sym.n
    z = np.empty_like(b)
    z[:, sym.pinv] = b                                  # cs_ipvec(pinv, b, z)
    sweep(sym, 9, Lx, z, seed)
    sweep(sym, 10, Ux, z, seed + 1)
    x = np.empty_like(b)
    x[:, sym.q] = z                                     # cs_ipvec(q, z, x)
    return x
