"""Seeded synthetic workloads of BASELINE.json's five configs.

numpy only; used by bench.py and the tests to build inputs.  All matrices are CSC with
int32 indices (sorted inside columns) and float64 values, like the reference's CscMat
(src/CSparse3/csc.py:107-123).

  C1  laplacian_2d(100)            n = 10,000
  C2  GridCase(118)                IEEE-118-shaped Newton-Raphson Jacobian
  C3  GridCase(2000).jacobian_batch  time-series batch on one pattern
  C4  GridCase(10000).outage_batch   N-1 contingency sweep on one pattern
  C5  laplacian_3d(100)            n = 1,000,000
"""
import numpy as np


def _coo_to_csc(m, n, rows, cols, vals):
    """Sorted CSC with duplicates summed (int32 / float64)."""
    key = cols.astype(np.int64) * m + rows.astype(np.int64)
    order = np.argsort(key, kind="stable")
    key = key[order]
    uniq, start = np.unique(key, return_index=True)
    data = np.add.reduceat(vals[order], start) if len(key) else np.zeros(0, dtype=vals.dtype)
    ci = (uniq % m).astype(np.int32)
    cj = (uniq // m).astype(np.int64)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, cj + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    return indptr, ci, data


def _tridiag_kron_terms(k, dims):
    """COO entries of the (2*dims+1)-point Laplacian on a k^dims grid."""
    N = k ** dims
    idx = np.arange(N, dtype=np.int64)
    rows = [idx]
    cols = [idx]
    vals = [np.full(N, 2.0 * dims)]
    for d in range(dims):
        stride = k ** d
        coord = (idx // stride) % k
        lo = idx[coord > 0]
        rows += [lo, lo - stride]
        cols += [lo - stride, lo]
        vals += [np.full(lo.size, -1.0), np.full(lo.size, -1.0)]
    return N, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def laplacian_2d(k=100):
    """C1: A = kron(I,T) + kron(T,I), T = tridiag(-1,2,-1).  -> (n, Ap, Ai, Ax)"""
    N, r, c, v = _tridiag_kron_terms(k, 2)
    Ap, Ai, Ax = _coo_to_csc(N, N, r, c, v)
    return N, Ap, Ai, Ax


def laplacian_3d(k=100):
    """C5: 7-point Laplacian on a k^3 grid.  -> (n, Ap, Ai, Ax)"""
    N, r, c, v = _tridiag_kron_terms(k, 3)
    Ap, Ai, Ax = _coo_to_csc(N, N, r, c, v)
    return N, Ap, Ai, Ax


def _bridges(n, f, t):
    """Boolean mask of bridge edges of the multigraph (iterative Tarjan low-link)."""
    m = len(f)
    adj_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(adj_ptr, f + 1, 1)
    np.add.at(adj_ptr, t + 1, 1)
    adj_ptr = np.cumsum(adj_ptr)
    fill = adj_ptr[:-1].copy()
    adj_v = np.empty(2 * m, dtype=np.int64)
    adj_e = np.empty(2 * m, dtype=np.int64)
    for e in range(m):
        a, b = f[e], t[e]
        adj_v[fill[a]] = b; adj_e[fill[a]] = e; fill[a] += 1
        adj_v[fill[b]] = a; adj_e[fill[b]] = e; fill[b] += 1
    disc = np.full(n, -1, dtype=np.int64)
    low = np.zeros(n, dtype=np.int64)
    bridge = np.zeros(m, dtype=bool)
    timer = 0
    for root in range(n):
        if disc[root] != -1:
            continue
        stack = [(root, -1, adj_ptr[root])]
        disc[root] = low[root] = timer; timer += 1
        while stack:
            v, pe, p = stack.pop()
            if p < adj_ptr[v + 1]:
                stack.append((v, pe, p + 1))
                w, e = adj_v[p], adj_e[p]
                if e == pe:
                    continue
                if disc[w] == -1:
                    disc[w] = low[w] = timer; timer += 1
                    stack.append((w, e, adj_ptr[w]))
                else:
                    low[v] = min(low[v], disc[w])
            elif stack:
                u = stack[-1][0]
                low[u] = min(low[u], low[v])
                if low[v] > disc[u]:
                    bridge[pe] = True
    return bridge


class GridCase:
    """Synthetic power grid (SURVEY 8d generator) and its polar Newton-Raphson Jacobian.

    Topology: spanning tree bus i <-> rng.integers(max(0,i-30), i) plus 0.4*N local chords
    a <-> a + rng.integers(1,30).  r~U(.01,.06), x~U(.05,.25), b_sh~U(0,.04); bus 0 slack,
    20 % PV, rest PQ.  Jacobian J = [[dP/dth, dP/dV],[dQ/dth, dQ/dV]] restricted to
    (pvpq, pq) -- the 2x2 block form pack_4_by_4 assembles (reference csc.py:588-606).
    """

    def __init__(self, n_bus, seed=0):
        rng = np.random.default_rng(seed)
        N = self.n_bus = int(n_bus)
        f = [np.arange(1, N, dtype=np.int64)]
        t = [np.array([rng.integers(max(0, i - 30), i) for i in range(1, N)], dtype=np.int64)]
        n_ch = int(0.4 * N)
        a = rng.integers(0, N, n_ch)
        b = np.minimum(a + rng.integers(1, 30, n_ch), N - 1)
        keep = a != b
        f.append(a[keep]); t.append(b[keep])
        self.f = np.concatenate(f)
        self.t = np.concatenate(t)
        nbr = self.n_branch = len(self.f)
        r = rng.uniform(0.01, 0.06, nbr)
        x = rng.uniform(0.05, 0.25, nbr)
        bsh = rng.uniform(0.0, 0.04, nbr)
        self.ys = 1.0 / (r + 1j * x)
        self.bsh = bsh
        types = np.full(N, 2)                       # 2 = PQ
        pv = rng.choice(np.arange(1, N), size=int(0.2 * N), replace=False)
        types[pv] = 1
        types[0] = 0
        self.pv = np.sort(pv)
        self.pq = np.where(types == 2)[0]
        self.pvpq = np.r_[self.pv, self.pq]
        self.seed = seed
        self._build_pattern()

    # Ybus pattern as COO of unique (i,k); per-branch contributions kept so that outages
    # (one branch admittance -> 0) are sparse value edits on a shared pattern.
    def _build_pattern(self):
        N, f, t = self.n_bus, self.f, self.t
        rows = np.concatenate([f, f, t, t])
        cols = np.concatenate([f, t, f, t])
        key = rows * N + cols
        diag_key = np.arange(N, dtype=np.int64) * (N + 1)
        uniq = np.unique(np.concatenate([key, diag_key]))
        self.yi = (uniq // N).astype(np.int64)
        self.yk = (uniq % N).astype(np.int64)
        self.nnz_y = len(uniq)
        self.br_slot = np.searchsorted(uniq, key).reshape(4, -1)      # ff, ft, tf, tt slots per branch
        self.y_rowstart = np.searchsorted(self.yi, np.arange(N))      # uniq is sorted by (i,k)
        self.y_isdiag = self.yi == self.yk
        # Jacobian pattern: entries (block, ybus entry) sorted by (col, row)
        npv_pq, npq = len(self.pvpq), len(self.pq)
        self.n = npv_pq + npq
        pos_th = np.full(N, -1, dtype=np.int64); pos_th[self.pvpq] = np.arange(npv_pq)
        pos_v = np.full(N, -1, dtype=np.int64); pos_v[self.pq] = np.arange(npq)
        jr, jc, jb, je = [], [], [], []
        ent = np.arange(self.nnz_y)
        for blk, (rp, cp, ro, co) in enumerate([(pos_th, pos_th, 0, 0), (pos_th, pos_v, 0, npv_pq),
                                                (pos_v, pos_th, npv_pq, 0), (pos_v, pos_v, npv_pq, npv_pq)]):
            sel = (rp[self.yi] >= 0) & (cp[self.yk] >= 0)
            jr.append(rp[self.yi[sel]] + ro); jc.append(cp[self.yk[sel]] + co)
            jb.append(np.full(sel.sum(), blk)); je.append(ent[sel])
        jr, jc, jb, je = map(np.concatenate, (jr, jc, jb, je))
        order = np.lexsort((jr, jc))
        self.Ai = jr[order].astype(np.int32)
        self.j_block = jb[order]
        self.j_ent = je[order]
        cnt = np.bincount(jc, minlength=self.n)
        self.Ap = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
        self.nnz = len(self.Ai)

    def ybus_values(self, out_branch=None):
        """Ybus values on the shared pattern; out_branch (array of branch ids, one per system) zeroes one
        branch per system -> shape [B, nnz_y] (or [nnz_y] when out_branch is None)."""
        yff = self.ys + 0.5j * self.bsh
        contrib = np.stack([yff, -self.ys, -self.ys, yff])               # [4, nbr]
        base = np.zeros(self.nnz_y, dtype=np.complex128)
        np.add.at(base, self.br_slot.ravel(), contrib.ravel())
        if out_branch is None:
            return base
        out_branch = np.asarray(out_branch)
        Y = np.broadcast_to(base, (len(out_branch), self.nnz_y)).copy()
        rows = np.arange(len(out_branch))
        for s in range(4):
            np.subtract.at(Y, (rows, self.br_slot[s, out_branch]), contrib[s, out_branch])
        return Y

    def voltages(self, seeds):
        """V[B, N] = U(.97,1.03) * exp(j*N(0,.05)), one seeded draw per system."""
        V = np.empty((len(seeds), self.n_bus), dtype=np.complex128)
        for r, s in enumerate(seeds):
            rng = np.random.default_rng(int(s))
            V[r] = rng.uniform(0.97, 1.03, self.n_bus) * np.exp(1j * rng.normal(0.0, 0.05, self.n_bus))
        return V

    def jacobian_values(self, V, Y=None):
        """Jacobian values on the shared pattern for voltages V[B,N] and Ybus values Y[B,nnz_y] or [nnz_y].
        dS/dVa = j*diag(V)*conj(diag(I) - Ybus*diag(V));  dS/dVm = diag(V)*conj(Ybus*diag(V/|V|)) + conj(diag(I))*diag(V/|V|)
        -> float64 [B, nnz]."""
        V = np.atleast_2d(V)
        if Y is None:
            Y = self.ybus_values()
        Y = np.broadcast_to(Y, (V.shape[0], self.nnz_y))
        Vi, Vk = V[:, self.yi], V[:, self.yk]
        YV = Y * Vk
        I = np.add.reduceat(YV, self.y_rowstart, axis=1)                  # I = Ybus * V   [B, N]
        Ii = I[:, self.yi] * self.y_isdiag
        Vn_k = Vk / np.abs(Vk)
        dVa = 1j * Vi * np.conj(Ii - YV)
        dVm = Vi * np.conj(Y * Vn_k) + np.conj(Ii) * Vn_k
        quad = np.stack([dVa.real, dVm.real, dVa.imag, dVm.imag])         # [4, B, nnz_y]
        return np.ascontiguousarray(quad[self.j_block, :, self.j_ent].T)  # [B, nnz]

    def base_jacobian(self):
        """-> (n, Ap, Ai, Ax) at the seed's base voltages (C2 single system)."""
        V = self.voltages([self.seed + 7])
        return self.n, self.Ap, self.Ai, self.jacobian_values(V)[0]

    def jacobian_batch(self, start, count):
        """C3 time-series: system k re-draws V with seed 1000+k; b_k = standard_normal(n).
        -> (Ax[count, nnz], b[count, n])"""
        ks = np.arange(start, start + count)
        Ax = self.jacobian_values(self.voltages(1000 + ks))
        b = np.stack([np.random.default_rng(int(2_000_000 + k)).standard_normal(self.n) for k in ks])
        return Ax, b

    def non_bridge_branches(self):
        if not hasattr(self, "_nb"):
            self._nb = np.where(~_bridges(self.n_bus, self.f, self.t))[0]
        return self._nb

    def outage_batch(self, start, count):
        """C4 N-1 sweep: system k = base case with non-bridge branch k out (admittance 0, entries kept
        as explicit values so the pattern is shared).  -> (Ax[count, nnz], b[count, n])"""
        nb = self.non_bridge_branches()
        ks = np.arange(start, min(start + count, len(nb)))
        V = np.broadcast_to(self.voltages([self.seed + 7]), (len(ks), self.n_bus))
        Ax = self.jacobian_values(V, self.ybus_values(nb[ks]))
        b = np.stack([np.random.default_rng(int(3_000_000 + k)).standard_normal(self.n) for k in ks])
        return Ax, b
