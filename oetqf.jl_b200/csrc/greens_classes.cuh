// greens_classes.cuh -- translation classes of (receiver, source) pairs for the hex8 builders (K3''/K4'').
//
// The hex8 kernel of a (receiver, cuboid) pair depends on the horizontal coordinates only through the offsets
// x_r - q_x and y_r - q_y (the half-space is invariant under horizontal translation); depths enter absolutely.
// The reference exploits the same invariance for its fault -> fault kernel (Toeplitz form, GF.jl:31-71) and
// evaluates every pair of the three mantle couplings (GF.jl:206-225, :262-290).  On the meshes the package
// builds -- Gmsh transfinite boxes (mesh.jl:95-130) and the equidistant fault (mesh.jl:39-56) -- a few thousand
// DISTINCT pairs stand for millions: the host sorts the pairs into classes, the closed form is evaluated once per
// class (the same per-pair code as K3/K4, on representative coordinates), and a copy kernel writes the dense
// shard at HBM speed.  Nothing is assumed about the mesh: classes are found numerically (coordinates equal to
// 1e-12 of the mesh extent are one value) and a mesh without enough structure simply yields as many classes as
// pairs, in which case the caller keeps the tiled kernels.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace oq {

// values equal within tol -> one class; index[i] = class of vals[i] (classes numbered in ascending order of value)
static inline int cluster_values(const std::vector<double>& vals, double tol, std::vector<int>& index)
{
    const size_t n = vals.size();
    std::vector<size_t> ord(n);
    for (size_t i = 0; i < n; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return vals[a] < vals[b]; });
    index.assign(n, 0);
    int ncls = 0;
    double anchor = 0.0;
    for (size_t k = 0; k < n; ++k) {
        const double v = vals[ord[k]];
        if (k == 0 || v - anchor > tol) { ++ncls; anchor = v; }    // a class never spans more than tol
        index[ord[k]] = ncls - 1;
    }
    return ncls;
}

static inline double span_of(const std::vector<double>& a, const std::vector<double>& b)
{
    double lo = 1e300, hi = -1e300;
    for (double v : a) { lo = std::min(lo, v); hi = std::max(hi, v); }
    for (double v : b) { lo = std::min(lo, v); hi = std::max(hi, v); }
    return hi > lo ? hi - lo : (hi == lo ? std::fabs(hi) : 0.0);
}

// acc[i] <- class of the tuple (acc[i], idx[i]); returns the number of distinct tuples
static inline int combine_classes(std::vector<int>& acc, int nacc, const std::vector<int>& idx, int nidx)
{
    (void)nacc;
    const size_t n = acc.size();
    std::vector<long long> key(n);
    for (size_t i = 0; i < n; ++i) key[i] = (long long)acc[i] * nidx + idx[i];
    std::vector<long long> uniq(key);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    for (size_t i = 0; i < n; ++i) acc[i] = (int)(std::lower_bound(uniq.begin(), uniq.end(), key[i]) - uniq.begin());
    return (int)uniq.size();
}

// Classes of (receiver, source) combinations along one group of coordinates: `t` is the coordinate that enters
// the kernel only through t_receiver - t_source; `attr` are class ids of everything else of that group that enters
// (sizes, depths).
struct AxisClasses {
    std::vector<int> rcls, scls;     // class of every receiver / source
    int nr = 0, ns = 0;
    std::vector<int> D;              // [nr * ns] -> pair class
    int n = 0;                       // number of pair classes
    std::vector<int> rep_r, rep_s;   // per pair class: a receiver / source (index into the input arrays) realising it

    // false: more than max_combos receiver-class x source-class combinations (no structure worth exploiting)
    bool build(const std::vector<double>& tr, const std::vector<int>& ar, int nar, const std::vector<double>& ts,
               const std::vector<int>& as, int nas, size_t max_combos)
    {
        const double tol = 1e-12 * span_of(tr, ts);
        std::vector<int> itr, its;
        const int ntr = cluster_values(tr, tol, itr), nts = cluster_values(ts, tol, its);
        rcls = itr; nr = combine_classes(rcls, ntr, ar, nar);
        scls = its; ns = combine_classes(scls, nts, as, nas);
        if ((size_t)nr * (size_t)ns > max_combos || (size_t)ntr * (size_t)nts > max_combos) return false;
        std::vector<int> rr(nr, -1), ss(ns, -1);          // first member of every class
        for (size_t i = 0; i < rcls.size(); ++i) if (rr[rcls[i]] < 0) rr[rcls[i]] = (int)i;
        for (size_t i = 0; i < scls.size(); ++i) if (ss[scls[i]] < 0) ss[scls[i]] = (int)i;
        // the offset of a combination depends on the VALUE classes of t only: cluster ntr x nts offsets, not nr x ns
        std::vector<double> vtr(ntr), vts(nts);
        {
            std::vector<char> seen_r(ntr, 0), seen_s(nts, 0);
            for (size_t i = 0; i < itr.size(); ++i) if (!seen_r[itr[i]]) { seen_r[itr[i]] = 1; vtr[itr[i]] = tr[i]; }
            for (size_t i = 0; i < its.size(); ++i) if (!seen_s[its[i]]) { seen_s[its[i]] = 1; vts[its[i]] = ts[i]; }
        }
        std::vector<double> off((size_t)ntr * nts);
        for (int a = 0; a < ntr; ++a)
            for (int b = 0; b < nts; ++b) off[(size_t)a * nts + b] = vtr[a] - vts[b];
        std::vector<int> offc;
        const int noff = cluster_values(off, tol, offc);
        // pair class of (a, b) = (offset class, receiver attributes, source attributes), numbered in order of first
        // appearance (row-major over receiver classes, then source classes)
        const size_t nc = (size_t)nr * ns;
        D.assign(nc, 0);
        rep_r.clear(); rep_s.clear();
        const unsigned long long space = (unsigned long long)noff * (unsigned long long)nar * (unsigned long long)nas;
        std::vector<int> direct;
        std::vector<std::pair<long long, int>> sorted_keys;
        const bool use_direct = space <= (1ull << 26);
        if (use_direct) direct.assign((size_t)space, -1);
        std::vector<long long> keys(use_direct ? 0 : nc);
        n = 0;
        for (int a = 0; a < nr; ++a) {
            const int ia = itr[rr[a]], ka = ar[rr[a]];
            for (int b = 0; b < ns; ++b) {
                const long long key = ((long long)offc[(size_t)ia * nts + its[ss[b]]] * nar + ka) * nas + as[ss[b]];
                if (use_direct) {
                    int& c = direct[(size_t)key];
                    if (c < 0) { c = n++; rep_r.push_back(rr[a]); rep_s.push_back(ss[b]); }
                    D[(size_t)a * ns + b] = c;
                } else {
                    keys[(size_t)a * ns + b] = key;
                }
            }
        }
        if (!use_direct) {      // huge key space: number the keys by sorting, then renumber in order of first appearance
            std::vector<long long> uniq(keys);
            std::sort(uniq.begin(), uniq.end());
            uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
            std::vector<int> first(uniq.size(), -1);
            for (size_t k = 0; k < nc; ++k) {
                const size_t u = std::lower_bound(uniq.begin(), uniq.end(), keys[k]) - uniq.begin();
                if (first[u] < 0) { first[u] = n++; rep_r.push_back(rr[k / ns]); rep_s.push_back(ss[k % ns]); }
                D[k] = first[u];
            }
        }
        return true;
    }

    // keep the receivers [begin, end) and the pair classes they use.  Classes and their representatives were
    // chosen over ALL receivers, so a row shard evaluates exactly the coordinates the full matrix would: shards
    // are bit-identical to the corresponding rows whatever the number of ranks.
    void restrict(int begin, int end)
    {
        std::vector<int> rmap(nr, -1), rlist, newr(end - begin);
        for (int r = begin; r < end; ++r) {
            const int a = rcls[r];
            if (rmap[a] < 0) { rmap[a] = (int)rlist.size(); rlist.push_back(a); }
            newr[r - begin] = rmap[a];
        }
        std::vector<int> cmap(n, -1), clist, newD(rlist.size() * (size_t)ns);
        for (size_t al = 0; al < rlist.size(); ++al)
            for (int b = 0; b < ns; ++b) {
                const int c = D[(size_t)rlist[al] * ns + b];
                if (cmap[c] < 0) { cmap[c] = (int)clist.size(); clist.push_back(c); }
                newD[al * ns + b] = cmap[c];
            }
        std::vector<int> nrr(clist.size()), nrs(clist.size());
        for (size_t k = 0; k < clist.size(); ++k) { nrr[k] = rep_r[clist[k]]; nrs[k] = rep_s[clist[k]]; }
        rcls.swap(newr); nr = (int)rlist.size(); D.swap(newD); n = (int)clist.size(); rep_r.swap(nrr); rep_s.swap(nrs);
    }
};

// Classes of the pairs (receiver, hex8 source cell): group 1 = x, group 2 = (y, z).
struct Hex8PairClasses {
    AxisClasses g1, g23;
    long long pairs = 0, classes = 0;
    bool worthwhile = false;      // decided on the WHOLE problem (>= 4 pairs per class) so that every row shard of a
                                  // matrix takes the same path as the full matrix (shards stay bit-identical to it)

    // receivers: horizontal position (rx, ry); everything else of a receiver that enters the kernel is summarised in
    // the class ids rax (with the x group: its x size) and rayz (with the (y,z) group: depth, y and z sizes)
    bool build(const OqHex8Mesh* ma, const std::vector<double>& rx, const std::vector<int>& rax, int nrax,
               const std::vector<double>& ry, const std::vector<int>& rayz, int nrayz)
    {
        const int ne = ma->n;
        std::vector<double> sdx(ma->dx, ma->dx + ne), sdy(ma->dy, ma->dy + ne), sdz(ma->dz, ma->dz + ne);
        std::vector<double> sqx(ma->qx, ma->qx + ne), sqy(ma->qy, ma->qy + ne), sqz(ma->qz, ma->qz + ne);
        std::vector<int> cdx, cdy, cdz, cqz;
        auto size_tol = [](const std::vector<double>& v) { double m = 0; for (double x : v) m = std::max(m, std::fabs(x)); return 1e-12 * m; };
        const int ndx = cluster_values(sdx, size_tol(sdx), cdx);
        const int ndy = cluster_values(sdy, size_tol(sdy), cdy);
        const int ndz = cluster_values(sdz, size_tol(sdz), cdz);
        const int nqz = cluster_values(sqz, 1e-12 * span_of(sqz, sqz), cqz);
        std::vector<int> syz(cdy);
        int nsyz = combine_classes(syz, ndy, cqz, nqz);
        nsyz = combine_classes(syz, nsyz, cdz, ndz);
        const size_t cap = (size_t)16 << 20;       // receiver-class x source-class combinations per group (host memory: ~30 B each)
        if (!g1.build(rx, rax, nrax, sqx, cdx, ndx, cap)) return false;
        if (!g23.build(ry, rayz, nrayz, sqy, syz, nsyz, cap)) return false;
        pairs = (long long)rx.size() * ne;
        classes = (long long)g1.n * g23.n;
        return true;
    }

    void restrict(int begin, int end)
    {
        worthwhile = (long long)g1.n * g23.n * 4 <= (long long)g1.rcls.size() * (long long)g1.scls.size();
        g1.restrict(begin, end); g23.restrict(begin, end);
        pairs = (long long)(end - begin) * (long long)g1.scls.size();
        classes = (long long)g1.n * g23.n;
    }
};

// classes of (receiver cell, source cell) pairs of the mantle -> mantle kernel (GF.jl:250-290): x group =
// (c_x - q_x, receiver dx, source dx), (y,z) group = (c_y - q_y, both dy, receiver c_z and dz, source q_z and dz);
// restricted to the receivers [e_begin, e_end).  Representatives are GLOBAL element indices.
static inline bool mantle_mantle_classes(const OqHex8Mesh* ma, int e_begin, int e_end, Hex8PairClasses& pc)
{
    const int ne = ma->n;
    std::vector<double> rx(ma->cx, ma->cx + ne), ry(ma->cy, ma->cy + ne), rz(ma->cz, ma->cz + ne);
    std::vector<double> rdx(ma->dx, ma->dx + ne), rdy(ma->dy, ma->dy + ne), rdz(ma->dz, ma->dz + ne);
    auto tol_of = [](const std::vector<double>& v) { double m = 0; for (double x : v) m = std::max(m, std::fabs(x)); return 1e-12 * m; };
    std::vector<int> rax, rayz, cz_, cdz_;
    const int nrax = cluster_values(rdx, tol_of(rdx), rax);
    int nrayz = cluster_values(rdy, tol_of(rdy), rayz);
    const int ncz = cluster_values(rz, 1e-12 * span_of(rz, rz), cz_);
    const int ncdz = cluster_values(rdz, tol_of(rdz), cdz_);
    nrayz = combine_classes(rayz, nrayz, cz_, ncz);
    nrayz = combine_classes(rayz, nrayz, cdz_, ncdz);
    if (!pc.build(ma, rx, rax, nrax, ry, rayz, nrayz)) return false;
    pc.restrict(e_begin, e_end);
    return true;
}

// classes of (fault cell, source cell) pairs of the mantle -> fault kernel (GF.jl:194-227), receivers = fault cells
// (vec index i + j*nx): x group = (x_f - q_x, dx), (y,z) group = (y_f - q_y, dy, z_f, q_z, dz); restricted to the
// cells [row_begin, row_end).  Representatives are GLOBAL cell indices.
static inline bool mantle_fault_classes(const OqHex8Mesh* ma, const OqFaultMesh* mf, int row_begin, int row_end,
                                        Hex8PairClasses& pc)
{
    const int nr = mf->nx * mf->nxi;
    std::vector<double> rx(nr), ry(nr), rz(nr);
    for (int fc = 0; fc < nr; ++fc) {
        rx[fc] = mf->x[fc % mf->nx]; ry[fc] = mf->y[fc / mf->nx]; rz[fc] = mf->z[fc / mf->nx];
    }
    std::vector<int> rax(nr, 0), rayz;
    const int nrz = cluster_values(rz, 1e-12 * span_of(rz, rz), rayz);
    if (!pc.build(ma, rx, rax, 1, ry, rayz, nrz)) return false;
    pc.restrict(row_begin, row_end);
    return true;
}

// ---- fault -> mantle (Okada, GF.jl:123-174) ------------------------------------------------------------------
// classes of tuples of doubles that are BITWISE equal (rows of `keys`, `len` doubles each)
static inline int exact_classes(const std::vector<double>& keys, size_t len, std::vector<int>& index)
{
    const size_t n = len ? keys.size() / len : 0;
    std::vector<size_t> ord(n);
    for (size_t i = 0; i < n; ++i) ord[i] = i;
    auto cmp = [&](size_t a, size_t b) { return memcmp(&keys[a * len], &keys[b * len], len * sizeof(double)); };
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return cmp(a, b) < 0; });
    index.assign(n, 0);
    int ncls = 0;
    for (size_t k = 0; k < n; ++k) {
        if (k == 0 || cmp(ord[k - 1], ord[k]) != 0) ++ncls;
        index[ord[k]] = ncls - 1;
    }
    return ncls;
}

// Classes of (receiver element, fault patch) pairs whose dc3d arguments are bitwise equal.  Strike direction: the
// differences x_w - (al + r*lrept) for every quadrature point w and periodic image r, formed exactly as the kernel
// forms them; (y, z): a receiver's (cy, dy, cz, dz) and the down-dip index of the patch (no invariance: all
// combinations are distinct).  Restricted to the receivers [e_begin, e_end); representatives are GLOBAL indices
// (element; patch i + j*nx).
static inline bool fault_mantle_classes(const OqFaultMesh* mf, const OqHex8Mesh* ma, const double* qc, int nq, int nrept,
                                        double lrept, int e_begin, int e_end, Hex8PairClasses& pc)
{
    const int ne = ma->n, nx = mf->nx, nxi = mf->nxi, nf = nx * nxi;
    AxisClasses& g1 = pc.g1;
    AxisClasses& g23 = pc.g23;
    {   // x group
        std::vector<double> rk(2 * (size_t)ne);
        for (int e = 0; e < ne; ++e) { rk[2 * e] = ma->cx[e]; rk[2 * e + 1] = ma->dx[e]; }
        g1.nr = exact_classes(rk, 2, g1.rcls);
        std::vector<int> rr(g1.nr, -1);
        for (int e = 0; e < ne; ++e) if (rr[g1.rcls[e]] < 0) rr[g1.rcls[e]] = e;
        g1.ns = nx;
        g1.scls.resize(nf);
        for (int j = 0; j < nf; ++j) g1.scls[j] = j % nx;
        const size_t len = 2 * (size_t)nq * (2 * nrept + 1);
        if ((size_t)g1.nr * nx * len > ((size_t)64 << 20)) return false;
        std::vector<double> keys((size_t)g1.nr * nx * len);
        for (int a = 0; a < g1.nr; ++a) {
            const double cx = ma->cx[rr[a]], hx = ma->dx[rr[a]] / 2;
            for (int q1 = 0; q1 < nx; ++q1) {
                double* k = &keys[((size_t)a * nx + q1) * len];
                for (int w = 0; w < nq; ++w) {
                    volatile double prod = qc[3 * w] * hx;          // product and sum rounded separately, as on the device
                    const double rx = cx + prod;
                    for (int r = -nrept; r <= nrept; ++r) {
                        volatile double jump = r * lrept;
                        const double a1 = mf->ax0[q1] + jump, a2 = mf->ax1[q1] + jump;
                        *k++ = rx - a1; *k++ = rx - a2;
                    }
                }
            }
        }
        g1.n = exact_classes(keys, len, g1.D);
        g1.rep_r.assign(g1.n, -1); g1.rep_s.assign(g1.n, -1);
        for (int a = 0; a < g1.nr; ++a)
            for (int q1 = 0; q1 < nx; ++q1) {
                const int c = g1.D[(size_t)a * nx + q1];
                if (g1.rep_r[c] < 0) { g1.rep_r[c] = rr[a]; g1.rep_s[c] = q1; }
            }
    }
    {   // (y, z) group
        std::vector<double> rk(4 * (size_t)ne);
        for (int e = 0; e < ne; ++e) { rk[4 * e] = ma->cy[e]; rk[4 * e + 1] = ma->dy[e]; rk[4 * e + 2] = ma->cz[e]; rk[4 * e + 3] = ma->dz[e]; }
        g23.nr = exact_classes(rk, 4, g23.rcls);
        std::vector<int> rr(g23.nr, -1);
        for (int e = 0; e < ne; ++e) if (rr[g23.rcls[e]] < 0) rr[g23.rcls[e]] = e;
        g23.ns = nxi;
        g23.scls.resize(nf);
        for (int j = 0; j < nf; ++j) g23.scls[j] = j / nx;
        if ((size_t)g23.nr * nxi > ((size_t)16 << 20)) return false;
        g23.n = g23.nr * nxi;
        g23.D.resize((size_t)g23.n); g23.rep_r.resize(g23.n); g23.rep_s.resize(g23.n);
        for (int b = 0; b < g23.nr; ++b)
            for (int q2 = 0; q2 < nxi; ++q2) {
                const int c = b * nxi + q2;
                g23.D[c] = c; g23.rep_r[c] = rr[b]; g23.rep_s[c] = q2 * nx;
            }
    }
    pc.restrict(e_begin, e_end);
    return true;
}

// ---- device side ------------------------------------------------------------------------------------------
struct ClassView {
    const int *rc1, *rc23;      // [nrows]  receiver classes
    const int *sc1, *sc23;      // [ne]     source classes
    const int *D1, *D23;        // [nr1*ns1], [nr23*ns23]
    int ns1, ns23, n1, n23;
};

struct DevPairClasses {
    DevBuf<int> rc1, rc23, sc1, sc23, D1, D23, rep_r1, rep_s1, rep_r23, rep_s23;
    ClassView v{};
    int upload(const Hex8PairClasses& c)
    {
        OQ_TRY(rc1.upload(c.g1.rcls.data(), c.g1.rcls.size())); OQ_TRY(rc23.upload(c.g23.rcls.data(), c.g23.rcls.size()));
        OQ_TRY(sc1.upload(c.g1.scls.data(), c.g1.scls.size())); OQ_TRY(sc23.upload(c.g23.scls.data(), c.g23.scls.size()));
        OQ_TRY(D1.upload(c.g1.D.data(), c.g1.D.size())); OQ_TRY(D23.upload(c.g23.D.data(), c.g23.D.size()));
        OQ_TRY(rep_r1.upload(c.g1.rep_r.data(), c.g1.rep_r.size())); OQ_TRY(rep_s1.upload(c.g1.rep_s.data(), c.g1.rep_s.size()));
        OQ_TRY(rep_r23.upload(c.g23.rep_r.data(), c.g23.rep_r.size())); OQ_TRY(rep_s23.upload(c.g23.rep_s.data(), c.g23.rep_s.size()));
        v.rc1 = rc1.p; v.rc23 = rc23.p; v.sc1 = sc1.p; v.sc23 = sc23.p; v.D1 = D1.p; v.D23 = D23.p;
        v.ns1 = c.g1.ns; v.ns23 = c.g23.ns; v.n1 = c.g1.n; v.n23 = c.g23.n;
        return 0;
    }
};

// G[(k*nrows + r), p*ne + i] = T[k*P + p][class of (r, i)]: the dense shard from the class table.  One row unit r
// per blockIdx.y step, sources across the threads (consecutive sources are consecutive columns AND, on a mesh
// numbered x-fastest, consecutive x classes: loads and stores of a warp are contiguous).  The table is small and
// re-read constantly (L2); the shard is written once and never read here (streaming stores).
template <int K, int P>
__global__ void __launch_bounds__(256)
expand_classes_kernel(const double* __restrict__ T, ClassView c, int nrows, int ne, size_t ld, double* __restrict__ G)
{
    const size_t tstride = (size_t)c.n1 * c.n23;
    for (int r = blockIdx.y; r < nrows; r += gridDim.y) {
        const int* d1 = c.D1 + (size_t)c.rc1[r] * c.ns1;
        const int* d23 = c.D23 + (size_t)c.rc23[r] * c.ns23;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += gridDim.x * blockDim.x) {
            const double* src = T + (size_t)__ldg(d23 + __ldg(c.sc23 + i)) * c.n1 + __ldg(d1 + __ldg(c.sc1 + i));
            double v[K * P];
#pragma unroll
            for (int m = 0; m < K * P; ++m) v[m] = __ldg(src + m * tstride);
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int p = 0; p < P; ++p)
                    __stcs(G + ((size_t)k * nrows + r) * ld + (size_t)p * ne + i, v[k * P + p]);
        }
        if (blockIdx.x == 0)                         // padding columns of the row unit
            for (size_t col = (size_t)P * ne + threadIdx.x; col < ld; col += blockDim.x)
#pragma unroll
                for (int k = 0; k < K; ++k) G[((size_t)k * nrows + r) * ld + col] = 0.0;
    }
}

}  // namespace oq
