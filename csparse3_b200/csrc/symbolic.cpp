// symbolic.cpp -- host symbolic phase (see symbolic.hpp).
#include "symbolic.hpp"
#include "program.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace csp3 {

static inline i32 flip(i32 i) { return -i - 2; }

// ---------------------------------------------------------------------------------------------------------
// pattern helpers
// ---------------------------------------------------------------------------------------------------------
Pattern transpose_pattern(i64 m, i64 n, const i32 *Ap, const i32 *Ai)
{
    Pattern T;
    T.m = n; T.n = m;
    T.p.assign((size_t)m + 1, 0);
    T.i.resize((size_t)std::max<i64>(Ap[n], 1));
    std::vector<i32> cursor((size_t)std::max<i64>(m, 1), 0);
    for (i32 e = 0; e < Ap[n]; ++e) cursor[Ai[e]]++;
    i32 run = 0;
    for (i64 r = 0; r < m; ++r) { T.p[r] = run; run += cursor[r]; cursor[r] = T.p[r]; }
    T.p[m] = run;
    for (i64 c = 0; c < n; ++c)
        for (i32 e = Ap[c]; e < Ap[c + 1]; ++e) T.i[cursor[Ai[e]]++] = (i32)c;
    return T;
}

// first-touch union of the columns of A and B (cs_add on patterns)
static Pattern union_pattern(i64 m, i64 n, const i32 *Ap, const i32 *Ai, const i32 *Bp, const i32 *Bi)
{
    Pattern C;
    C.m = m; C.n = n;
    C.p.resize((size_t)n + 1);
    C.i.reserve((size_t)Ap[n] + Bp[n] + 1);
    std::vector<i32> seen((size_t)std::max<i64>(m, 1), 0);
    for (i64 c = 0; c < n; ++c) {
        C.p[c] = (i32)C.i.size();
        const i32 stamp = (i32)c + 1;
        for (i32 e = Ap[c]; e < Ap[c + 1]; ++e)
            if (seen[Ai[e]] < stamp) { seen[Ai[e]] = stamp; C.i.push_back(Ai[e]); }
        for (i32 e = Bp[c]; e < Bp[c + 1]; ++e)
            if (seen[Bi[e]] < stamp) { seen[Bi[e]] = stamp; C.i.push_back(Bi[e]); }
    }
    C.p[n] = (i32)C.i.size();
    return C;
}

// first-touch pattern of A*B (cs_multiply on patterns); A is Am x An, B is An x Bn
static Pattern product_pattern(i64 Am, const i32 *Ap, const i32 *Ai, i64 Bn, const i32 *Bp, const i32 *Bi)
{
    Pattern C;
    C.m = Am; C.n = Bn;
    C.p.resize((size_t)Bn + 1);
    std::vector<i32> seen((size_t)std::max<i64>(Am, 1), 0);
    for (i64 c = 0; c < Bn; ++c) {
        C.p[c] = (i32)C.i.size();
        const i32 stamp = (i32)c + 1;
        for (i32 eb = Bp[c]; eb < Bp[c + 1]; ++eb) {
            const i32 k = Bi[eb];
            for (i32 ea = Ap[k]; ea < Ap[k + 1]; ++ea)
                if (seen[Ai[ea]] < stamp) { seen[Ai[ea]] = stamp; C.i.push_back(Ai[ea]); }
        }
    }
    C.p[Bn] = (i32)C.i.size();
    return C;
}

// ---------------------------------------------------------------------------------------------------------
// approximate minimum degree on a quotient graph (CSparse cs_amd semantics, tie-breaks included)
// ---------------------------------------------------------------------------------------------------------
namespace {

struct QuotientGraph {
    i32 n;
    std::vector<i32> ptr;      // Cp: start of each node/element list; flip()-encoded parent once absorbed
    std::vector<i32> adj;      // Ci: list storage with elbow room
    i64 used, cap;             // cnz, nzmax
    std::vector<i32> len, nv, next, head, elen, degree, w, hhead, last;
    i32 mark = 0, lemax = 0, mindeg = 0, nel = 0, dense;

    // state of the pivot step in flight
    i32 k = -1, elenk = 0, nvk = 0, dk = 0, pk1 = 0, pk2 = 0;

    QuotientGraph(Pattern &C, i32 n_, i32 dense_) : n(n_), dense(dense_)
    {
        ptr.swap(C.p);
        adj.swap(C.i);
        used = ptr[n];
        cap = used + used / 5 + 2 * (i64)n;
        adj.resize((size_t)std::max<i64>(cap, 1));
        const size_t sz = (size_t)n + 1;
        len.assign(sz, 0); nv.assign(sz, 1); next.assign(sz, -1); head.assign(sz, -1); elen.assign(sz, 0);
        degree.assign(sz, 0); w.assign(sz, 1); hhead.assign(sz, -1); last.assign(sz, -1);
        for (i32 v = 0; v < n; ++v) { len[v] = ptr[v + 1] - ptr[v]; degree[v] = len[v]; }
        len[n] = 0; degree[n] = 0;
        mark = clear_marks(0);
        elen[n] = -2; ptr[n] = -1; w[n] = 0;
    }

    i32 clear_marks(i32 m)
    {
        if (m < 2 || m + lemax < 0) {
            for (i32 v = 0; v < n; ++v) if (w[v] != 0) w[v] = 1;
            m = 2;
        }
        return m;
    }

    void degree_list_push(i32 v, i32 d)
    {
        if (head[d] != -1) last[head[d]] = v;
        next[v] = head[d];
        head[d] = v;
    }

    void seed_degree_lists()
    {
        for (i32 v = 0; v < n; ++v) {
            const i32 d = degree[v];
            if (d == 0) { elen[v] = -2; nel++; ptr[v] = -1; w[v] = 0; }
            else if (d > dense) { nv[v] = 0; elen[v] = -1; nel++; ptr[v] = flip(n); nv[n]++; }
            else degree_list_push(v, d);
        }
    }

    void pick_pivot()
    {
        for (k = -1; mindeg < n && (k = head[mindeg]) == -1; mindeg++) {}
        if (next[k] != -1) last[next[k]] = -1;
        head[mindeg] = next[k];
        elenk = elen[k];
        nvk = nv[k];
        nel += nvk;
    }

    void compact_if_needed()
    {
        if (!(elenk > 0 && used + mindeg >= cap)) return;
        for (i32 j = 0; j < n; ++j) {
            const i32 p = ptr[j];
            if (p >= 0) { ptr[j] = adj[p]; adj[p] = flip(j); }
        }
        i32 dst = 0;
        for (i32 src = 0; src < used;) {
            const i32 j = flip(adj[src++]);
            if (j >= 0) {
                adj[dst] = ptr[j];
                ptr[j] = dst++;
                for (i32 c = 0; c < len[j] - 1; ++c) adj[dst++] = adj[src++];
            }
        }
        used = dst;
    }

    void form_element()
    {
        dk = 0;
        nv[k] = -nvk;
        i32 p = ptr[k];
        pk1 = (elenk == 0) ? p : (i32)used;
        pk2 = pk1;
        for (i32 step = 1; step <= elenk + 1; ++step) {
            i32 e, pj, ln;
            if (step > elenk) { e = k; pj = p; ln = len[k] - elenk; }
            else { e = adj[p++]; pj = ptr[e]; ln = len[e]; }
            for (i32 c = 0; c < ln; ++c) {
                const i32 v = adj[pj++], nvv = nv[v];
                if (nvv <= 0) continue;
                dk += nvv;
                nv[v] = -nvv;
                adj[pk2++] = v;
                if (next[v] != -1) last[next[v]] = last[v];
                if (last[v] != -1) next[last[v]] = next[v];
                else head[degree[v]] = next[v];
            }
            if (e != k) { ptr[e] = flip(k); w[e] = 0; }
        }
        if (elenk != 0) used = pk2;
        degree[k] = dk;
        ptr[k] = pk1;
        len[k] = pk2 - pk1;
        elen[k] = -2;
    }

    void scan_set_differences()
    {
        mark = clear_marks(mark);
        for (i32 pk = pk1; pk < pk2; ++pk) {
            const i32 v = adj[pk], eln = elen[v];
            if (eln <= 0) continue;
            const i32 nvv = -nv[v], wnv = mark - nvv;
            for (i32 p = ptr[v]; p <= ptr[v] + eln - 1; ++p) {
                const i32 e = adj[p];
                if (w[e] >= mark) w[e] -= nvv;
                else if (w[e] != 0) w[e] = degree[e] + wnv;
            }
        }
    }

    void update_degrees()
    {
        for (i32 pk = pk1; pk < pk2; ++pk) {
            const i32 v = adj[pk];
            const i32 p1 = ptr[v], p2 = p1 + elen[v] - 1;
            i32 pn = p1, d = 0;
            i64 h = 0;
            for (i32 p = p1; p <= p2; ++p) {
                const i32 e = adj[p];
                if (w[e] == 0) continue;
                const i32 dext = w[e] - mark;
                if (dext > 0) { d += dext; adj[pn++] = e; h += e; }
                else { ptr[e] = flip(k); w[e] = 0; }     // aggressive absorption
            }
            elen[v] = pn - p1 + 1;
            const i32 p3 = pn, p4 = p1 + len[v];
            for (i32 p = p2 + 1; p < p4; ++p) {
                const i32 u = adj[p], nvu = nv[u];
                if (nvu <= 0) continue;
                d += nvu;
                adj[pn++] = u;
                h += u;
            }
            if (d == 0) {                                    // mass elimination
                ptr[v] = flip(k);
                const i32 nvv = -nv[v];
                dk -= nvv; nvk += nvv; nel += nvv;
                nv[v] = 0; elen[v] = -1;
            } else {
                degree[v] = std::min(degree[v], d);
                adj[pn] = adj[p3];
                adj[p3] = adj[p1];
                adj[p1] = k;
                len[v] = pn - p1 + 1;
                h = ((h < 0) ? -h : h) % n;
                next[v] = hhead[h];
                hhead[h] = v;
                last[v] = (i32)h;
            }
        }
        degree[k] = dk;
        lemax = std::max(lemax, dk);
        mark = clear_marks(mark + lemax);
    }

    void merge_indistinguishable()
    {
        for (i32 pk = pk1; pk < pk2; ++pk) {
            i32 v = adj[pk];
            if (nv[v] >= 0) continue;
            const i32 h = last[v];
            v = hhead[h];
            hhead[h] = -1;
            for (; v != -1 && next[v] != -1; v = next[v], mark++) {
                const i32 ln = len[v], eln = elen[v];
                for (i32 p = ptr[v] + 1; p <= ptr[v] + ln - 1; ++p) w[adj[p]] = mark;
                i32 prev = v;
                for (i32 u = next[v]; u != -1;) {
                    bool same = (len[u] == ln) && (elen[u] == eln);
                    for (i32 p = ptr[u] + 1; same && p <= ptr[u] + ln - 1; ++p)
                        if (w[adj[p]] != mark) same = false;
                    if (same) {
                        ptr[u] = flip(v);
                        nv[v] += nv[u];
                        nv[u] = 0;
                        elen[u] = -1;
                        u = next[u];
                        next[prev] = u;
                    } else {
                        prev = u;
                        u = next[u];
                    }
                }
            }
        }
    }

    void close_element()
    {
        i32 dst = pk1;
        for (i32 pk = pk1; pk < pk2; ++pk) {
            const i32 v = adj[pk], nvv = -nv[v];
            if (nvv <= 0) continue;
            nv[v] = nvv;
            i32 d = degree[v] + dk - nvv;
            d = std::min(d, n - nel - nvv);
            degree_list_push(v, d);
            last[v] = -1;
            mindeg = std::min(mindeg, d);
            degree[v] = d;
            adj[dst++] = v;
        }
        nv[k] = nvk;
        if ((len[k] = dst - pk1) == 0) { ptr[k] = -1; w[k] = 0; }
        if (elenk != 0) used = dst;
    }

    std::vector<i32> assembly_postorder()
    {
        std::vector<i32> perm((size_t)n + 1, 0);
        for (i32 v = 0; v < n; ++v) ptr[v] = flip(ptr[v]);
        for (i32 v = 0; v <= n; ++v) head[v] = -1;
        for (i32 v = n; v >= 0; --v) {
            if (nv[v] > 0) continue;
            next[v] = head[ptr[v]];
            head[ptr[v]] = v;
        }
        for (i32 e = n; e >= 0; --e) {
            if (nv[e] <= 0) continue;
            if (ptr[e] != -1) { next[e] = head[ptr[e]]; head[ptr[e]] = e; }
        }
        i32 out = 0;
        std::vector<i32> stack((size_t)n + 1);
        for (i32 r = 0; r <= n; ++r) {
            if (ptr[r] != -1) continue;
            i32 top = 0;
            stack[0] = r;
            while (top >= 0) {
                const i32 p = stack[top], c = head[p];
                if (c == -1) { --top; perm[out++] = p; }
                else { head[p] = next[c]; stack[++top] = c; }
            }
        }
        perm.resize((size_t)n);
        return perm;
    }
};

}  // namespace

std::vector<i32> amd_order(i64 order, i64 m_, i64 n_, const i32 *Ap, const i32 *Ai)
{
    const i32 m = (i32)m_, n = (i32)n_;
    std::vector<i32> q((size_t)n);
    if (order <= 0 || order > 3 || n == 0) {
        for (i32 c = 0; c < n; ++c) q[c] = c;
        return q;
    }
    Pattern AT = transpose_pattern(m, n, Ap, Ai);
    i32 dense = (i32)std::max<double>(16.0, 10.0 * std::sqrt((double)n));
    dense = std::min(n - 2, dense);
    Pattern C;
    if (order == 1 && n == m) {
        C = union_pattern(m, n, Ap, Ai, AT.p.data(), AT.i.data());
    } else if (order == 2) {
        i32 keep = 0;
        for (i32 r = 0; r < m; ++r) {                      // drop dense rows of A (= dense columns of A')
            i32 e = AT.p[r];
            AT.p[r] = keep;
            if (AT.p[r + 1] - e > dense) continue;
            for (; e < AT.p[r + 1]; ++e) AT.i[keep++] = AT.i[e];
        }
        AT.p[m] = keep;
        Pattern S = transpose_pattern(n, m, AT.p.data(), AT.i.data());   // S = A without dense rows (m x n)
        C = product_pattern(n, AT.p.data(), AT.i.data(), n, S.p.data(), S.i.data());
    } else {
        C = product_pattern(n, AT.p.data(), AT.i.data(), n, Ap, Ai);
    }
    // remove the diagonal, keep order
    {
        i32 dst = 0;
        for (i32 c = 0; c < n; ++c) {
            i32 e = C.p[c];
            C.p[c] = dst;
            for (; e < C.p[c + 1]; ++e)
                if (C.i[e] != c) C.i[dst++] = C.i[e];
        }
        C.p[n] = dst;
    }
    QuotientGraph G(C, n, dense);
    G.seed_degree_lists();
    while (G.nel < n) {
        G.pick_pivot();
        G.compact_if_needed();
        G.form_element();
        G.scan_set_differences();
        G.update_degrees();
        G.merge_indistinguishable();
        G.close_element();
    }
    return G.assembly_postorder();
}

// ---------------------------------------------------------------------------------------------------------
// elimination tree / postorder (CSparse cs_etree, cs_post)
// ---------------------------------------------------------------------------------------------------------
std::vector<i32> etree(i64 m, i64 n, const i32 *Ap, const i32 *Ai, bool ata)
{
    std::vector<i32> parent((size_t)n, -1), ancestor((size_t)n, -1), prev;
    if (ata) prev.assign((size_t)m, -1);
    for (i64 c = 0; c < n; ++c) {
        for (i32 e = Ap[c]; e < Ap[c + 1]; ++e) {
            i32 r = ata ? prev[Ai[e]] : Ai[e];
            while (r != -1 && r < c) {
                const i32 up = ancestor[r];
                ancestor[r] = (i32)c;
                if (up == -1) parent[r] = (i32)c;
                r = up;
            }
            if (ata) prev[Ai[e]] = (i32)c;
        }
    }
    return parent;
}

std::vector<i32> postorder(i64 n, const i32 *parent)
{
    std::vector<i32> head((size_t)n, -1), next((size_t)n, -1), stack((size_t)n), post((size_t)n);
    for (i64 v = n - 1; v >= 0; --v) {
        if (parent[v] == -1) continue;
        next[v] = head[parent[v]];
        head[parent[v]] = (i32)v;
    }
    i32 out = 0;
    for (i64 r = 0; r < n; ++r) {
        if (parent[r] != -1) continue;
        i32 top = 0;
        stack[0] = (i32)r;
        while (top >= 0) {
            const i32 p = stack[top], c = head[p];
            if (c == -1) { --top; post[out++] = p; }
            else { head[p] = next[c]; stack[++top] = c; }
        }
    }
    return post;
}

// ---------------------------------------------------------------------------------------------------------
// first factorisation: left-looking Gilbert-Peierls with threshold partial pivoting (CSparse cs_lu)
// ---------------------------------------------------------------------------------------------------------
namespace {

// Non-recursive DFS over the graph of L restricted to pivotal rows (CSparse cs_dfs / cs_reach).
struct ReachFinder {
    const std::vector<i32> &Lp, &Li, &pinv;
    std::vector<i32> order;     // xi: output stack filled from the top (size n)
    std::vector<i32> frame, resume;
    std::vector<char> visited;
    i32 n;
    ReachFinder(i32 n_, const std::vector<i32> &Lp_, const std::vector<i32> &Li_, const std::vector<i32> &pinv_)
        : Lp(Lp_), Li(Li_), pinv(pinv_), order((size_t)n_), frame((size_t)n_), resume((size_t)n_),
          visited((size_t)n_, 0), n(n_) {}

    i32 visit(i32 start, i32 top)
    {
        i32 depth = 0;
        frame[0] = start;
        while (depth >= 0) {
            const i32 r = frame[depth];
            const i32 col = pinv[r];
            if (!visited[r]) {
                visited[r] = 1;
                resume[depth] = (col < 0) ? 0 : Lp[col];
            }
            const i32 stop = (col < 0) ? 0 : Lp[col + 1];
            bool descended = false;
            for (i32 e = resume[depth]; e < stop; ++e) {
                const i32 child = Li[e];
                if (visited[child]) continue;
                resume[depth] = e;
                frame[++depth] = child;
                descended = true;
                break;
            }
            if (!descended) { --depth; order[--top] = r; }
        }
        return top;
    }

    // topological order of Reach(A(:,col)) in order[top..n-1]
    i32 reach(const i32 *Ap, const i32 *Ai, i32 col)
    {
        i32 top = n;
        for (i32 e = Ap[col]; e < Ap[col + 1]; ++e)
            if (!visited[Ai[e]]) top = visit(Ai[e], top);
        for (i32 t = top; t < n; ++t) visited[order[t]] = 0;
        return top;
    }
};

}  // namespace

int lu_factor(i64 n_, const i32 *Ap, const i32 *Ai, const double *Ax, const i32 *q, double tol, Factor &F)
{
    const i32 n = (i32)n_;
    F.Lp.assign((size_t)n + 1, 0);
    F.Up.assign((size_t)n + 1, 0);
    F.pinv.assign((size_t)n, -1);
    F.Li.clear(); F.Lx.clear(); F.Ui.clear(); F.Ux.clear();
    const size_t guess = 4 * (size_t)Ap[n] + (size_t)n;
    F.Li.reserve(guess); F.Lx.reserve(guess); F.Ui.reserve(guess); F.Ux.reserve(guess);
    std::vector<double> x((size_t)std::max(n, 1), 0.0);
    ReachFinder rf(n, F.Lp, F.Li, F.pinv);
    for (i32 k = 0; k < n; ++k) {
        F.Lp[k] = (i32)F.Li.size();
        F.Up[k] = (i32)F.Ui.size();
        const i32 col = q ? q[k] : k;
        const i32 top = rf.reach(Ap, Ai, col);
        const i32 *xi = rf.order.data();
        for (i32 t = top; t < n; ++t) x[xi[t]] = 0.0;
        for (i32 e = Ap[col]; e < Ap[col + 1]; ++e) x[Ai[e]] = Ax[e];
        // sparse triangular solve x = L \ A(:,col) over the reach
        for (i32 t = top; t < n; ++t) {
            const i32 r = xi[t], J = F.pinv[r];
            if (J < 0) continue;
            x[r] /= F.Lx[F.Lp[J]];
            const double xr = x[r];
            for (i32 e = F.Lp[J] + 1; e < F.Lp[J + 1]; ++e) x[F.Li[e]] -= F.Lx[e] * xr;
        }
        // pivot search; pivotal rows go to U
        i32 ipiv = -1;
        double best = -1.0;
        for (i32 t = top; t < n; ++t) {
            const i32 r = xi[t];
            if (F.pinv[r] < 0) {
                const double a = std::fabs(x[r]);
                if (a > best) { best = a; ipiv = r; }
            } else {
                F.Ui.push_back(F.pinv[r]);
                F.Ux.push_back(x[r]);
            }
        }
        if (ipiv == -1 || best <= 0.0) return k + 1;
        if (F.pinv[col] < 0 && std::fabs(x[col]) >= best * tol) ipiv = col;
        const double pivot = x[ipiv];
        F.Ui.push_back(k);
        F.Ux.push_back(pivot);
        F.pinv[ipiv] = k;
        F.Li.push_back(ipiv);
        F.Lx.push_back(1.0);
        for (i32 t = top; t < n; ++t) {
            const i32 r = xi[t];
            if (F.pinv[r] < 0) { F.Li.push_back(r); F.Lx.push_back(x[r] / pivot); }
            x[r] = 0.0;
        }
    }
    F.Lp[n] = (i32)F.Li.size();
    F.Up[n] = (i32)F.Ui.size();
    for (auto &r : F.Li) r = F.pinv[r];
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// level sets
// ---------------------------------------------------------------------------------------------------------
LevelSet build_levels(i64 n, const std::vector<i32> &Gp, const std::vector<i32> &Gi, int kind)
{
    LevelSet L;
    L.level.assign((size_t)n, 0);
    if (kind == 0) {
        for (i64 k = 0; k < n; ++k) {
            i32 lv = 0;
            for (i32 e = Gp[k]; e < Gp[k + 1]; ++e)
                if (Gi[e] < k) lv = std::max(lv, L.level[Gi[e]] + 1);
            L.level[k] = lv;
        }
    } else if (kind == 1) {
        for (i64 j = 0; j < n; ++j)
            for (i32 e = Gp[j]; e < Gp[j + 1]; ++e)
                if (Gi[e] > j) L.level[Gi[e]] = std::max(L.level[Gi[e]], L.level[j] + 1);
    } else {
        for (i64 j = n - 1; j >= 0; --j)
            for (i32 e = Gp[j]; e < Gp[j + 1]; ++e)
                if (Gi[e] < j) L.level[Gi[e]] = std::max(L.level[Gi[e]], L.level[j] + 1);
    }
    i32 nlev = 0;
    for (i64 k = 0; k < n; ++k) nlev = std::max(nlev, L.level[k] + 1);
    L.lptr.assign((size_t)nlev + 1, 0);
    for (i64 k = 0; k < n; ++k) L.lptr[L.level[k] + 1]++;
    for (i32 l = 0; l < nlev; ++l) L.lptr[l + 1] += L.lptr[l];
    std::vector<i32> cursor(L.lptr.begin(), L.lptr.end() - 1);
    L.order.resize((size_t)n);
    for (i64 k = 0; k < n; ++k) L.order[cursor[L.level[k]]++] = (i32)k;
    return L;
}

// ---------------------------------------------------------------------------------------------------------
// device schedule
// ---------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------
// program compiler (formats in program.hpp)
// ---------------------------------------------------------------------------------------------------------
namespace {

struct Emitter {
    std::vector<uint8_t> &out;
    size_t max_record = 0, rec_start = 0;
    explicit Emitter(std::vector<uint8_t> &o) : out(o) {}
    void i32v(i32 v) { const uint8_t *b = (const uint8_t *)&v; out.insert(out.end(), b, b + 4); }
    void u16v(i32 v) { const uint16_t w = (uint16_t)v; const uint8_t *b = (const uint8_t *)&w; out.insert(out.end(), b, b + 2); }
    void pad8() { while (out.size() & 7) out.push_back(0); }
    void begin() { rec_start = out.size(); }
    void end() { pad8(); max_record = std::max(max_record, out.size() - rec_start); }
    void finish(Program &P)
    {
        i32 stage = 512;
        while ((size_t)stage < max_record + 16) stage *= 2;
        while (out.size() % (size_t)stage) out.push_back(0);
        out.insert(out.end(), (size_t)stage, 0);            // one guard stage: look-ahead header reads stay in bounds
        P.stage = stage;
    }
};

void compile_programs(i64 n_, const std::vector<i32> &q, const Factor &F, Schedule &S)
{
    const i32 n = (i32)n_;
    // ---- refactor: column header + A list + prefetch directives are ONE record, every pair is its own record
    {
        S.rf_prog = Program();
        Emitter E(S.rf_prog.bytes);
        for (i32 k = 0; k < n; ++k) {
            const ColDesc &cd = S.cols[k];
            const i32 kp = k + kPfCols;
            const i32 pf_cnt = kp < n ? S.cols[kp].a_cnt : 0;
            E.begin();
            E.i32v(cd.up); E.i32v(cd.lp);
            E.u16v(cd.ucnt); E.u16v(cd.lcnt); E.u16v(cd.a_cnt); E.u16v(cd.pair_cnt);
            // far-back sources of column k + kPfMissCols (older than the recent-L ring)
            std::vector<std::pair<i32, i32>> far;
            const i32 km = k + kPfMissCols;
            if (km < n) {
                const ColDesc &cm = S.cols[km];
                for (i32 pi = cm.pair_ptr; pi < cm.pair_ptr + cm.pair_cnt; ++pi)
                    if (S.pairs[pi].lstart < cm.lp - kCompileWindow && S.pairs[pi].llen > 0)
                        far.emplace_back(S.pairs[pi].lstart, S.pairs[pi].llen);
            }
            E.u16v(pf_cnt); E.u16v((i32)far.size()); E.i32v(0);
            for (i32 t = 0; t < cd.a_cnt; ++t) E.i32v(S.a_src[cd.a_ptr + t]);
            for (i32 t = 0; t < cd.a_cnt; ++t) E.u16v(S.a_off[cd.a_ptr + t]);
            E.pad8();
            for (i32 t = 0; t < pf_cnt; ++t) E.i32v(S.a_src[S.cols[kp].a_ptr + t]);
            E.pad8();
            for (auto &f : far) { E.i32v(f.first); E.u16v(f.second); E.u16v(0); }
            E.end();
            for (i32 pi = cd.pair_ptr; pi < cd.pair_ptr + cd.pair_cnt; ++pi) {
                const PairDesc &pd = S.pairs[pi];
                E.begin();
                E.i32v(pd.lstart); E.u16v(pd.moff); E.u16v(pd.llen);
                for (i32 t = 0; t < pd.llen; ++t) E.u16v(S.upd_map[(size_t)pd.mapstart + t]);
                E.end();
            }
        }
        E.finish(S.rf_prog);
    }
    // ---- sweeps
    auto sweep = [&](const SolveStream &T, bool lower, Program &P) {
        P = Program();
        Emitter E(P.bytes);
        // forward: right-hand sides come from b (original row prow[i]); results are parked at x-position q[j].
        // backward: right-hand sides are those parked values (position q[i]); results go to q[j].
        auto xpos = [&](i32 row) { return q.empty() ? row : q[row]; };
        auto rhs_index = [&](i32 row) { return lower ? S.prow[row] : xpos(row); };
        for (i32 step = 0; step < n; ++step) {
            const i32 j = lower ? step : n - 1 - step;
            const SolveCol &c = T.cols[j];
            const i32 len = c.len_alloc & 0xffff, nalloc = (i32)((uint32_t)c.len_alloc >> 16);
            const i32 stepp = step + kPfCols;
            // right-hand sides needed kPfCols columns from now: that column's allocations and, when its own
            // row was never touched, its own right-hand side
            std::vector<i32> pf;
            if (stepp < n) {
                const i32 jp = lower ? stepp : n - 1 - stepp;
                const SolveCol &cp = T.cols[jp];
                const i32 na = (i32)((uint32_t)cp.len_alloc >> 16);
                for (i32 t = 0; t < na; ++t) pf.push_back(rhs_index(T.alloc_row[cp.alloc_ptr + t]));
                if (cp.slot < 0) pf.push_back(rhs_index(jp));
            }
            const i32 pf_cnt = (i32)pf.size();
            E.begin();
            E.i32v(c.start); E.i32v(rhs_index(j));
            E.u16v(c.slot); E.u16v(len); E.u16v(nalloc); E.u16v(pf_cnt);
            E.i32v(xpos(j)); E.i32v(0);
            for (i32 t = 0; t < len; ++t) E.u16v(T.slot[(size_t)c.start + t]);
            E.pad8();
            for (i32 t = 0; t < nalloc; ++t) E.i32v(rhs_index(T.alloc_row[c.alloc_ptr + t]));
            for (i32 t = 0; t < nalloc; ++t) E.u16v(T.alloc_slot[c.alloc_ptr + t]);
            E.pad8();
            for (i32 t = 0; t < pf_cnt; ++t) E.i32v(pf[t]);
            E.end();
        }
        E.finish(P);
    };
    sweep(S.ls, true, S.ls_prog);
    sweep(S.us, false, S.us_prog);
    // ---- row-oriented backward sweep
    {
        const std::vector<i32> &Up = F.Up, &Ui = F.Ui;
        auto xpos = [&](i32 row) { return q.empty() ? row : q[row]; };
        std::vector<i32> last_row((size_t)std::max(n, 1), -1);       // smallest row that uses x_j (-1: none)
        for (i32 j = 0; j < n; ++j)
            for (i32 e = Up[j]; e < Up[j + 1] - 1; ++e)
                if (last_row[j] < 0 || Ui[e] < last_row[j]) last_row[j] = Ui[e];
        struct RowRec { i32 diagpos, xpos, slot_out; std::vector<i32> pos, slot; };
        std::vector<RowRec> recs((size_t)n);
        std::vector<i32> slot_of((size_t)std::max(n, 1), -1), free_list;
        i32 next_slot = 0;
        S.ur_max_len = 0;
        for (i32 i = n - 1; i >= 0; --i) {
            RowRec &R = recs[n - 1 - i];
            R.diagpos = Up[i + 1] - 1;
            R.xpos = xpos(i);
            for (i32 t = S.urow_ptr[i + 1] - 1; t >= S.urow_ptr[i]; --t) {            // descending column
                R.pos.push_back(S.urow_pos[t]);
                R.slot.push_back(slot_of[S.urow_col[t]]);
            }
            S.ur_max_len = std::max(S.ur_max_len, (i32)R.pos.size());
            for (i32 t = S.urow_ptr[i]; t < S.urow_ptr[i + 1]; ++t) {
                const i32 j = S.urow_col[t];
                if (last_row[j] == i) { free_list.push_back(slot_of[j]); slot_of[j] = -1; }
            }
            R.slot_out = -1;
            if (last_row[i] >= 0) {                                                   // some later row needs x_i
                if (!free_list.empty()) { R.slot_out = free_list.back(); free_list.pop_back(); }
                else R.slot_out = next_slot++;
                slot_of[i] = R.slot_out;
            }
        }
        S.ur_nslots = next_slot;
        S.ur_prog = Program();
        Emitter E(S.ur_prog.bytes);
        for (i32 step = 0; step < n; ++step) {
            const RowRec &R = recs[step];
            const RowRec *Pf = step + kPfCols < n ? &recs[step + kPfCols] : nullptr;
            const i32 pf_cnt = Pf ? (i32)Pf->pos.size() + 1 : 0;
            E.begin();
            E.i32v(R.diagpos); E.i32v(R.xpos);
            E.u16v(R.slot_out); E.u16v((i32)R.pos.size()); E.u16v(pf_cnt); E.u16v(0);
            E.i32v(Pf ? Pf->xpos : -1); E.i32v(0);
            for (size_t t = 0; t < R.pos.size(); ++t) { E.i32v(R.pos[t]); E.u16v(R.slot[t]); E.u16v(0); }
            if (Pf) { for (i32 v : Pf->pos) E.i32v(v); E.i32v(Pf->diagpos); }
            E.end();
        }
        E.finish(S.ur_prog);
    }
}

}  // namespace

bool build_schedule(i64 n_, const i32 *Ap, const i32 *Ai, const std::vector<i32> &q, const Factor &F,
                    Schedule &S, const char **why)
{
    const i32 n = (i32)n_;
    const std::vector<i32> &Lp = F.Lp, &Li = F.Li, &Up = F.Up, &Ui = F.Ui;
    S = Schedule();
    S.cols.resize((size_t)n);
    S.a_src.resize((size_t)Ap[n]);
    S.a_off.resize((size_t)Ap[n]);
    S.pairs.reserve((size_t)Up[n]);
    std::vector<i32> slot_of_row((size_t)std::max(n, 1), -1);   // row (pivot numbering) -> accumulator slot
    std::vector<i32> entry_of_slot(65536, -1);
    std::vector<std::pair<i32, i32>> by_pivot;                   // (j, slot) of U(:,k) off-diagonals
    i64 map_total = 0;
    i32 a_cursor = 0;
    for (i32 k = 0; k < n; ++k) {
        ColDesc &cd = S.cols[k];
        cd.up = Up[k]; cd.lp = Lp[k];
        cd.ucnt = Up[k + 1] - Up[k];
        cd.lcnt = Lp[k + 1] - Lp[k];
        const i32 len = cd.ucnt + cd.lcnt - 1;
        if (len > 65535) { *why = "a factor column has more than 65535 entries"; return false; }
        S.max_col_len = std::max(S.max_col_len, len);
        for (i32 t = 0; t < cd.ucnt; ++t) slot_of_row[Ui[Up[k] + t]] = t;
        for (i32 t = 1; t < cd.lcnt; ++t) slot_of_row[Li[Lp[k] + t]] = cd.ucnt + t - 1;
        // scatter list of A(:,q[k])
        const i32 col = q.empty() ? k : q[k];
        cd.a_ptr = a_cursor;
        cd.a_cnt = Ap[col + 1] - Ap[col];
        for (i32 e = Ap[col]; e < Ap[col + 1]; ++e) {
            const i32 slot = slot_of_row[F.pinv[Ai[e]]];
            if (slot < 0) { *why = "A entry outside the factor pattern"; return false; }
            // duplicate (row, col) entries: cs_lu's scatter x[Ai[p]] = Ax[p] keeps the LAST one
            if (entry_of_slot[slot] >= cd.a_ptr) { S.a_src[entry_of_slot[slot]] = e; continue; }
            entry_of_slot[slot] = a_cursor;
            S.a_src[a_cursor] = e;
            S.a_off[a_cursor] = (uint16_t)slot;
            ++a_cursor;
        }
        cd.a_cnt = a_cursor - cd.a_ptr;
        for (i32 t = cd.a_ptr; t < a_cursor; ++t) entry_of_slot[S.a_off[t]] = -1;
        // update pairs in the STORED order of U(:,k): the topological (reach) order the first factorisation
        // used, so every accumulator slot sees its updates in exactly the order of CSparse cs_lu / the oracle
        by_pivot.clear();
        for (i32 t = 0; t < cd.ucnt - 1; ++t) by_pivot.emplace_back(Ui[Up[k] + t], t);
        cd.pair_ptr = (i32)S.pairs.size();
        cd.pair_cnt = (i32)by_pivot.size();
        for (auto &jt : by_pivot) {
            const i32 j = jt.first;
            PairDesc pd;
            pd.moff = jt.second;
            pd.lstart = Lp[j] + 1;
            pd.llen = Lp[j + 1] - Lp[j] - 1;
            if (map_total + pd.llen > INT32_MAX) { *why = "more than 2^31-1 update slots"; return false; }
            pd.mapstart = (i32)map_total;
            map_total += pd.llen;
            S.pairs.push_back(pd);
        }
        // fill the map for this column's pairs
        S.upd_map.resize((size_t)map_total);
        for (i32 pi = cd.pair_ptr; pi < cd.pair_ptr + cd.pair_cnt; ++pi) {
            const PairDesc &pd = S.pairs[pi];
            for (i32 t = 0; t < pd.llen; ++t) {
                const i32 slot = slot_of_row[Li[pd.lstart + t]];
                if (slot < 0) { *why = "update target outside the factor pattern"; return false; }
                S.upd_map[(size_t)pd.mapstart + t] = (uint16_t)slot;
            }
        }
        for (i32 t = 0; t < cd.ucnt; ++t) slot_of_row[Ui[Up[k] + t]] = -1;
        for (i32 t = 1; t < cd.lcnt; ++t) slot_of_row[Li[Lp[k] + t]] = -1;
    }
    S.flops = 2 * map_total;
    S.lev_refactor = build_levels(n, Up, Ui, 0);
    S.lev_lsolve = build_levels(n, Lp, Li, 1);
    S.lev_usolve = build_levels(n, Up, Ui, 2);
    // row views of strict L and strict U
    auto row_view = [n](const std::vector<i32> &Gp, const std::vector<i32> &Gi, bool lower,
                        std::vector<i32> &rp, std::vector<i32> &rc, std::vector<i32> &rpos) {
        rp.assign((size_t)n + 1, 0);
        for (i32 c = 0; c < n; ++c)
            for (i32 e = Gp[c]; e < Gp[c + 1]; ++e)
                if (lower ? (Gi[e] > c) : (Gi[e] < c)) rp[Gi[e] + 1]++;
        for (i32 r = 0; r < n; ++r) rp[r + 1] += rp[r];
        rc.resize((size_t)rp[n]); rpos.resize((size_t)rp[n]);
        std::vector<i32> cur(rp.begin(), rp.end() - 1);
        for (i32 c = 0; c < n; ++c)
            for (i32 e = Gp[c]; e < Gp[c + 1]; ++e)
                if (lower ? (Gi[e] > c) : (Gi[e] < c)) { rc[cur[Gi[e]]] = c; rpos[cur[Gi[e]]++] = e; }
    };
    row_view(Lp, Li, true, S.lrow_ptr, S.lrow_col, S.lrow_pos);
    row_view(Up, Ui, false, S.urow_ptr, S.urow_col, S.urow_pos);
    // streaming solves
    S.prow.assign((size_t)n, 0);
    for (i32 r = 0; r < n; ++r) S.prow[F.pinv[r]] = r;
    auto stream = [n](const std::vector<i32> &Gp, const std::vector<i32> &Gi, bool lower, SolveStream &T,
                      const char **why_) -> bool {
        T = SolveStream();
        T.cols.resize((size_t)n);
        T.slot.assign(Gi.size(), 0);
        std::vector<i32> slot_of((size_t)std::max(n, 1), -1), free_list;
        i32 next_slot = 0;
        for (i32 step = 0; step < n; ++step) {
            const i32 j = lower ? step : n - 1 - step;
            // strictly lower part of L(:,j) is [Lp[j]+1, Lp[j+1]); strictly upper part of U(:,j) is [Up[j], Up[j+1]-1)
            const i32 beg = lower ? Gp[j] + 1 : Gp[j], end = lower ? Gp[j + 1] : Gp[j + 1] - 1;
            SolveCol &c = T.cols[j];
            c.slot = slot_of[j];
            c.start = beg;
            c.alloc_ptr = (i32)T.alloc_row.size();
            if (slot_of[j] >= 0) { free_list.push_back(slot_of[j]); slot_of[j] = -1; }   // y[j] is read before any update below
            i32 nalloc = 0;
            for (i32 p = beg; p < end; ++p) {
                const i32 i = Gi[p];
                if (slot_of[i] < 0) {
                    i32 sl;
                    if (!free_list.empty()) { sl = free_list.back(); free_list.pop_back(); }
                    else sl = next_slot++;
                    slot_of[i] = sl;
                    T.alloc_row.push_back(i);
                    T.alloc_slot.push_back((uint16_t)sl);
                    ++nalloc;
                }
                T.slot[p] = (uint16_t)slot_of[i];
            }
            if (end - beg > 65535 || next_slot > 65535) { *why_ = "solve wavefront exceeds 65535 rows"; return false; }
            c.len_alloc = (end - beg) | (nalloc << 16);
        }
        T.nslots = next_slot;
        return true;
    };
    if (!stream(Lp, Li, true, S.ls, why) || !stream(Up, Ui, false, S.us, why)) return false;
    compile_programs(n, q, F, S);
    return true;
}

}  // namespace csp3
