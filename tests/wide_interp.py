"""CPU interpreter of the compiled "wide" refactor program (csparse3_b200/csrc/program.hpp), test infrastructure.

Executes the byte stream record by record exactly as lu_wide.cu does -- same accumulator slots, same L cache /
landing area, same fetch timing -- on a handful of systems at once (numpy vectors over the batch), so the host
compiler (wide_program.cpp) can be validated against the oracle without a GPU.  cp.async is modelled at its two
extremes: the destination is poisoned (NaN) when a fetch is issued, because the data may land at any moment from
then on, and the data only becomes readable kWideLookahead records later (or at once for an immediate fetch).
"""
import ctypes as C
import struct

import numpy as np

from csparse3_b200 import _lib

LOOKAHEAD = 4
AREGS = 6
COL_HEADER = 48
CHUNK_HEADER = 16
PROG_STAGES = 8


def get_program(sym, which=3):
    geo = (C.c_int64 * 8)()
    size = _lib.lib().csp3_lu_get_program(sym._h, which, None, 0, geo)
    if size < 0:
        return None, None
    buf = np.empty(size, dtype=np.uint8)
    _lib.lib().csp3_lu_get_program(sym._h, which, buf.ctypes.data_as(C.c_void_p), size, geo)
    return buf.tobytes(), [int(v) for v in geo]


def run_refactor(sym, Ax):
    """Ax: [B, nnz] -> (Lx [B, lnz], Ux [B, unz], stats) by interpreting the wide program."""
    prog, geo = get_program(sym, 3)
    assert prog is not None, "wide program not available"
    stage, width, acc_slots, ring, land, nrec, smem, groups = geo[:8]
    EB = 8 * width
    cover = AREGS * groups
    ring_bytes = PROG_STAGES * stage
    B = Ax.shape[0]
    n, lnz, unz = sym.n, sym.lnz, sym.unz
    nval = acc_slots + 16 + ring + land               # value area, in entries: acc, pivot table, lsrc
    val = np.full((nval, B), np.nan)
    Lg = np.full((lnz, B), np.nan)
    Ug = np.full((unz, B), np.nan)
    Lg[sym.Lp[:-1]] = 1.0                              # unit diagonal: never written by the kernel, but a fetch of several
    AxT = np.ascontiguousarray(Ax.T)                   # adjacent L columns reads across it (the value is not used)
    pending = []                       # (ready_record, dst entry, data)
    rec_no = 0
    p = 0
    cur_stage = 0
    an = None
    fail = np.zeros(B, dtype=np.int64)
    stats = dict(fetches=0, immediates=0, records=0, ops=0)

    def entry_of(byte_off):
        assert byte_off % EB == 0
        return byte_off // EB

    def advance(p0, nbytes, flags):
        """checks the span of the record at p0 and the stage / wrap flags; returns the offset of the next record"""
        nonlocal cur_stage
        assert p0 % 16 == 0
        assert p0 // ring_bytes == (p0 + nbytes - 1) // ring_bytes, "record straddles the program ring"
        assert nbytes <= stage, "record larger than a stage"
        adv = (flags >> 1) & 3
        assert p0 // stage == cur_stage + adv, "stage flags wrong"
        cur_stage += adv
        nxt = p0 + nbytes
        if flags & 8:
            nxt = (nxt + ring_bytes - 1) // ring_bytes * ring_bytes
        return nxt

    def issue(units, dst16, src16, immediate):
        nonlocal pending
        if units == 0:
            return
        flen, fdst, fsrc = entry_of(units * 16), entry_of(dst16 * 16), entry_of(src16 * 16)
        assert fdst >= acc_slots + 16 + ring and fdst + flen <= nval, "fetch outside the landing area"
        data = Lg[fsrc:fsrc + flen].copy()
        assert not np.isnan(data).any(), "fetch of an L column that is not final yet"
        val[fdst:fdst + flen] = np.nan
        pending.append((rec_no if immediate else rec_no + LOOKAHEAD, fdst, data))
        stats["fetches"] += 1
        stats["immediates"] += int(immediate)

    def land_ready():
        nonlocal pending
        keep = []
        for ready, dst, data in pending:
            if ready <= rec_no:
                val[dst:dst + len(data)] = data
            else:
                keep.append((ready, dst, data))
        pending = keep

    GROUP_COLS = 8
    table0 = acc_slots                                 # pivot / reciprocal entries live right behind the accumulator
    lsrc0 = acc_slots + 2 * GROUP_COLS
    cap = 2 * groups
    cols_done = 0
    first = True
    while cols_done < n or first:
        (ncols, nslots, a_cnt, chunk_cnt, fin_cnt, an_cnt, fdst16, funits, fsrc16, npf) = struct.unpack_from("<HHHHHHHHiH", prog, p + 8)
        flags = prog[p + 34]
        cdesc = np.frombuffer(prog, dtype=np.int32, count=2 * GROUP_COLS, offset=p + COL_HEADER).reshape(GROUP_COLS, 2)
        pfd = np.frombuffer(prog, dtype=np.int32, count=2 * GROUP_COLS, offset=p + COL_HEADER + 8 * GROUP_COLS).reshape(GROUP_COLS, 2)
        lists = COL_HEADER + 16 * GROUP_COLS
        so = p + lists
        no = p + ((lists + 2 * a_cnt + 3) & ~3)
        over = max(0, a_cnt - cover)
        nbytes = (((lists + 2 * a_cnt + 3) & ~3) + 4 * (an_cnt + over) + 15) & ~15
        slots = [entry_of(int(v)) for v in np.frombuffer(prog, dtype=np.uint16, count=a_cnt, offset=so)]
        srcs = np.frombuffer(prog, dtype=np.int32, count=an_cnt + over, offset=no)
        p = advance(p, nbytes, flags)
        issue(funits, fdst16, fsrc16, False)
        land_ready()
        if first:
            assert ncols == 0 and a_cnt == 0 and chunk_cnt == 0 and fin_cnt == 0
            val[:acc_slots] = 0.0                      # the kernel clears the accumulator once
        first = False
        assert nslots <= acc_slots and ncols <= GROUP_COLS
        assert (val[:acc_slots] == 0.0).all(), "accumulator not clean at the start of a group"
        for t in range(a_cnt):
            assert slots[t] < nslots
            if t < cover:
                val[slots[t]] = an[t]
            else:
                val[slots[t]] = AxT[srcs[an_cnt + t - cover]]
        an = [AxT[srcs[t]].copy() for t in range(an_cnt)]
        for c in range(npf):
            if pfd[c, 0] >= 0:
                assert 0 <= pfd[c, 0] and pfd[c, 0] + pfd[c, 1] <= sym.nnz
        rec_no += 1
        for _ in range(chunk_cnt):
            fsrc16, fdst16, funits, flags = struct.unpack_from("<iHHH", prog, p)
            nbytes = CHUNK_HEADER + 8 * cap
            ent = np.frombuffer(prog, dtype=np.uint16, count=4 * cap, offset=p + CHUNK_HEADER).reshape(cap, 4)
            p = advance(p, nbytes, flags)
            issue(funits, fdst16, fsrc16, bool(flags & 1))
            land_ready()
            ok = ent[:, 3] != 0
            assert ok.any() and set(np.unique(ent[:, 3]).tolist()) <= {0, 1}
            for e in range(groups):                    # one multiplier load per lane group (entries e and e + groups)
                if ok[e + groups]:
                    assert ok[e] and ent[e, 1] == ent[e + groups, 1], "entries of a lane group have different multipliers"
            src = np.array([entry_of(int(v)) for v in ent[ok, 0]])
            mul = np.array([entry_of(int(v)) for v in ent[ok, 1]])
            tgt = np.array([entry_of(int(v)) for v in ent[ok, 2]])
            assert (src >= lsrc0).all() and (mul < nslots).all() and (tgt < nslots).all()
            assert len(set(tgt.tolist())) == len(tgt), "two operations of a chunk share a target"
            assert not (set(mul.tolist()) & set(tgt.tolist())), "a multiplier is modified inside its chunk"
            lv = val[src]
            assert not np.isnan(lv).any(), "chunk reads a source that has not landed (record %d)" % rec_no
            val[tgt] = val[tgt] - lv * val[mul]         # all loads, then all stores
            stats["ops"] += int(ok.sum())
            rec_no += 1
        # pivots of the group's columns (reciprocal pass), then the finalisation records
        pivots = []
        for c in range(ncols):
            col1, w = int(cdesc[c, 0]), int(cdesc[c, 1]) & 0xffff
            assert col1 > 0
            pv = val[entry_of(w)].copy()
            pivots.append(pv)
            bad = ~((np.abs(pv) > 0) & np.isfinite(pv))
            fail = np.where(bad & ((fail == 0) | (fail > col1)), col1, fail)
        nfin = 0
        for _ in range(fin_cnt):
            fsrc16, fdst16, funits, flags = struct.unpack_from("<iHHH", prog, p)
            nbytes = CHUNK_HEADER + 8 * cap
            fe = np.frombuffer(prog, dtype=np.int32, count=2 * cap, offset=p + CHUNK_HEADER).reshape(cap, 2)
            p = advance(p, nbytes, flags)
            issue(funits, fdst16, fsrc16, False)
            land_ready()
            for u in range(cap):
                gout, w = int(fe[u, 0]) & 0xffffffff, int(fe[u, 1]) & 0xffffffff
                slot_off, cache_off = w & 0xffff, w >> 16
                if slot_off == 0xffff:
                    continue
                assert bool(gout & 0x80000000) == bool(flags & 16), "finalisation record mixes U and L entries"
                sl = entry_of(slot_off)
                x = val[sl].copy()
                val[sl] = 0.0
                pos = gout & 0x0fffffff
                if gout & 0x80000000:
                    cidx = (gout >> 28) & 7
                    assert cidx < ncols
                    with np.errstate(all="ignore"):
                        q = x / pivots[cidx]
                    Lg[pos] = q
                    if cache_off != 0xffff:
                        ce = entry_of(cache_off)
                        assert lsrc0 <= ce < lsrc0 + ring
                        val[ce] = q
                else:
                    Ug[pos] = x
                nfin += 1
            rec_no += 1
        assert nfin == nslots
        cols_done += ncols
    Lg[sym.Lp[:-1]] = 1.0                              # the unit diagonal is implicit in the workspace layout
    stats["records"] = rec_no
    assert rec_no == nrec
    return np.ascontiguousarray(Lg.T), np.ascontiguousarray(Ug.T), fail, stats


SWEEP_LOOKAHEAD = 3
SWEEP_STAGES = 6


def _run_sweep(sym, which, Fx, zin):
    """Interprets one sweep program.  Fx: [B, lnz|unz] factor values, zin: [n, B] right-hand sides -> zout [n, B]."""
    prog, geo = get_program(sym, which)
    assert prog is not None, "wide sweep program not available"
    stage, width, nslots, landing, _, nrec, smem, E = geo[:8]
    EB = 8 * width
    LA, NL = SWEEP_LOOKAHEAD, SWEEP_LOOKAHEAD + 1
    SET = landing // NL
    EH = E // 2                                    # load / finalisation entries per record
    assert SET in (2 * E, 2 * E + EH)
    RB = 16 + 26 * E
    ring_bytes = SWEEP_STAGES * stage
    n = sym.n
    B = zin.shape[1]
    FT = np.ascontiguousarray(Fx.T)
    slots = np.full((nslots, B), np.nan)
    land = np.full((landing, B), np.nan)
    zout = np.full((n, B), np.nan)
    pending = []                                   # (ready record, kind, dst, data)
    p = 0
    cur_stage = 0
    nops = 0
    for r in range(nrec):
        flags, = struct.unpack_from("<H", prog, p)
        assert p % 8 == 0 and p // ring_bytes == (p + RB - 1) // ring_bytes, "record straddles the program ring"
        adv = (flags >> 1) & 3
        assert p // stage == cur_stage + adv, "stage flags wrong"
        cur_stage += adv
        loads = np.frombuffer(prog, dtype=np.int32, count=2 * EH, offset=p + 16).reshape(EH, 2)
        pf = np.frombuffer(prog, dtype=np.int32, count=2 * E, offset=p + 16 + 8 * EH)
        pfd = np.frombuffer(prog, dtype=np.int32, count=EH, offset=p + 16 + 8 * EH + 8 * E)
        fins = np.frombuffer(prog, dtype=np.int32, count=2 * EH, offset=p + 16 + 12 * EH + 8 * E).reshape(EH, 2)
        upds = np.frombuffer(prog, dtype=np.uint16, count=4 * E, offset=p + 16 + 20 * EH + 8 * E).reshape(2 * E, 2)
        cyc, pset = r % NL, (r + LA) % NL
        # issue: loads into slots, gathers into the landing set of record r + LA (poisoned until they land)
        for e in range(EH):
            gidx, w = int(loads[e, 0]), int(loads[e, 1])
            if gidx >= 0:
                so = (w & 0xffff) * 16
                assert so % EB == 0 and so // EB < nslots
                slots[so // EB] = np.nan
                pending.append((r + LA, 0, so // EB, zin[gidx].copy()))
        for u in range(2 * E):
            if pf[u] >= 0:
                land[pset * SET + u] = np.nan
                pending.append((r + LA, 1, pset * SET + u, FT[pf[u]].copy()))
        for e in range(EH):
            if pfd[e] >= 0:
                assert SET == 2 * E + EH
                land[pset * SET + 2 * E + e] = np.nan
                pending.append((r + LA, 1, pset * SET + 2 * E + e, FT[pfd[e]].copy()))
        keep = []
        for ready, kind, dst, data in pending:
            if ready <= r:
                (slots if kind == 0 else land)[dst] = data
            else:
                keep.append((ready, kind, dst, data))
        pending = keep
        # finalisations
        for e in range(EH):
            out, w = int(fins[e, 0]), int(fins[e, 1])
            if out >= 0:
                so, div = (w & 0xffff) * 16 // EB, (w >> 16) & 0xffff
                v = slots[so].copy()
                assert not np.isnan(v).any(), "finalisation of a row that has not landed (record %d)" % r
                if div:
                    d = land[cyc * SET + 2 * E + e]
                    assert not np.isnan(d).any()
                    v = v / d
                    slots[so] = v
                zout[out] = v
        # updates: all loads, then all stores
        ok = upds[:, 1] != 0xffff
        if ok.any():
            idx = np.nonzero(ok)[0]
            mul = upds[idx, 0].astype(np.int64) * 16 // EB
            tgt = upds[idx, 1].astype(np.int64) * 16 // EB
            assert len(set(tgt.tolist())) == len(tgt) and not (set(tgt.tolist()) & set(mul.tolist()))
            lv = land[cyc * SET + idx]
            assert not np.isnan(lv).any() and not np.isnan(slots[mul]).any() and not np.isnan(slots[tgt]).any(), \
                "update reads a value that has not landed (record %d)" % r
            slots[tgt] = slots[tgt] - lv * slots[mul]
            nops += len(idx)
        p += RB
        if flags & 8:
            p = (p + ring_bytes - 1) // ring_bytes * ring_bytes
    assert not np.isnan(zout).any()
    return zout, nops


def run_solve(sym, Lx, Ux, b):
    """b: [B, n] -> x [B, n] by interpreting the forward and backward sweep programs on factors Lx, Ux."""
    n = sym.n
    z1 = np.empty((n, b.shape[0]))
    z1[sym.pinv] = b.T                           # y = P b
    z2, nf = _run_sweep(sym, 4, Lx, z1)
    z3, nb = _run_sweep(sym, 5, Ux, z2)
    x = np.empty((n, b.shape[0]))
    x[sym.q] = z3                                # x[q[i]] = z[i]
    assert nf == sym.lnz - n and nb == sym.unz - n
    return np.ascontiguousarray(x.T)
