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
#include <climits>
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
        // local class ids keep the ORDER of the global ones (the class-form matvec walks a receiver's sources in
        // ascending class order: the order, hence every rounded sum, must not depend on the shard)
        {
            std::vector<int> sorted(clist);
            std::sort(sorted.begin(), sorted.end());
            for (size_t k = 0; k < sorted.size(); ++k) cmap[sorted[k]] = (int)k;
            for (size_t al = 0; al < rlist.size(); ++al)
                for (int b = 0; b < ns; ++b) newD[al * ns + b] = cmap[D[(size_t)rlist[al] * ns + b]];
            clist.swap(sorted);
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
// The table is addressed as T[class * cstride + m * mstride]: (1, n1*n23) for the builders' layout T[m][c23][c1],
// (ts, 1) for the class-major layout of a class-form operand (ClassOperand::Tm).
template <int K, int P>
__global__ void __launch_bounds__(256)
expand_classes_kernel(const double* __restrict__ T, ClassView c, int nrows, int ne, size_t ld, double* __restrict__ G,
                      size_t cstride = 1, size_t mstride = 0)
{
    const size_t tstride = mstride ? mstride : (size_t)c.n1 * c.n23;
    for (int r = blockIdx.y; r < nrows; r += gridDim.y) {
        const int* d1 = c.D1 + (size_t)c.rc1[r] * c.ns1;
        const int* d23 = c.D23 + (size_t)c.rc23[r] * c.ns23;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += gridDim.x * blockDim.x) {
            const double* src = T + ((size_t)__ldg(d23 + __ldg(c.sc23 + i)) * c.n1 + __ldg(d1 + __ldg(c.sc1 + i))) * cstride;
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

// ---- class form kept for the RHS (classmat.cuh) -------------------------------------------------------------
// Tm[c][m] (class-major, ts doubles per class) from the builders' T[m][c]
__global__ void __launch_bounds__(256)
class_table_transpose_kernel(const double* __restrict__ T, size_t ncls, int KP, int ts, double* __restrict__ Tm)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncls * ts) return;
    const size_t c = t / ts;
    const int m = (int)(t % ts);
    Tm[t] = m < KP ? T[(size_t)m * ncls + c] : 0.0;
}

// Td[c23][o][k][c] (38 doubles per block): the table in OFFSET order for the sliding-window kernel, zero where there
// is no such offset / residue.  idx = o - npad + NS indexes dcls[slot][idx] (see OffsetPlan); src blocks have `ts_src`
// doubles: mode 0 [k*6+c], mode 1 (1x6) [c] with slot = k, mode 2 (6x1) [k] with slot = c.
__global__ void __launch_bounds__(256)
class_table_offset_kernel(const double* __restrict__ Tm, int ts_src, const int* __restrict__ dcls, int nd, int mode, int Q,
                          size_t n23, int n1, int noff, int shift, double* __restrict__ Td)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n23 * noff * 38) return;
    const int e = (int)(t % 38);
    const size_t co = t / 38;
    const int o = (int)(co % noff);
    const size_t c23 = co / noff;
    double v = 0.0;
    const int idx = o - shift;
    if (e < 36 && idx >= 0 && idx < nd) {
        const int k = e / 6, c = e - k * 6;
        const int slot = mode == 1 ? k : mode == 2 ? c : 0;
        if (slot < Q || mode == 0) {
            const int cls = dcls[(size_t)slot * nd + idx];
            if (cls >= 0) v = Tm[(c23 * n1 + cls) * ts_src + (mode == 0 ? e : mode == 1 ? c : k)];
        }
    }
    Td[t] = v;
}

// The class form of a shard: table (class-major copy), class maps and the work list of class_matvec_kernel.
// `pc` is restricted to the shard's receivers; `nr` local receiver units, `ns` source units.

// ---- offset plan of the sliding-window kernel (class_matvec_diag_kernel) ---------------------------------------
// The kernel multiplies 6x6 blocks indexed by the OFFSET between a coarse receiver position m and a coarse source
// position j.  Three operand shapes map onto it:
//   mode 0  mantle -> mantle (6x6): one grid, m = j = x position; block = the class block [k][p]
//   mode 1  mantle -> fault  (1x6): sources on the coarse grid, receivers on a finer one: fine position = Q*m + rho;
//           block row k = residue rho of the receiver, columns = the 6 unit strains
//   mode 2  fault -> mantle  (6x1): receivers on the coarse grid, sources fine: position = Q*j + rho;
//           block rows = the 6 stress components, column c = residue rho of the source
// (up to 6 residues).  Everything is found numerically from integer positions on the common fine grid; `false` means the
// x classes are not a function of (slot, m - j) or the positions are not what the window needs, and the operand keeps the
// general kernel.
struct OffsetPlan {
    int mode = 0, Q = 1;                  // residues in use
    int MR = 0, NS = 0;                   // coarse receiver positions of the shard, coarse source positions
    std::vector<int> dcls;                // [6][MR + NS - 1]: class of (slot, m' - j' + NS - 1), -1: none
    std::vector<int> drow, dbeg, dcnt, dm0;   // runs of <= blk*8 consecutive coarse receiver positions of one (y,z) class
    std::vector<int> out_map;             // [entries][6]: flat index into y of (entry = coarse position of a run, row k), -1: none
    std::vector<int> xmap;                // [ns23][npad][6]: flat index into x of (source group, coarse position, column), -1: none
};

static inline bool floor_divmod(long long t, long long q, long long& m, long long& r)
{
    m = t / q; r = t - m * q;
    if (r < 0) { r += q; --m; }
    return true;
}

// rpos / spos: integer x positions (common fine grid) of the local receivers / all sources; c0, qstep: origin and step
// of the coarse grid in the same units (mode 0: qstep = 1)
static inline bool plan_offsets(const Hex8PairClasses& pc, int mode, const int* rpos, const int* spos, int nr, int ns,
                                long long c0, long long qstep, int K, int blk, int npad_for, OffsetPlan& pl)
{
    const AxisClasses& g1 = pc.g1;
    if (nr <= 0 || ns <= 0 || qstep <= 0) return false;
    pl.mode = mode;
    // coarse index and residue of every receiver / source
    std::vector<long long> mr(nr), rr(nr), js(ns), rs(ns);
    for (int r = 0; r < nr; ++r) floor_divmod((long long)rpos[r] - c0, qstep, mr[r], rr[r]);
    for (int s2 = 0; s2 < ns; ++s2) floor_divmod((long long)spos[s2] - c0, qstep, js[s2], rs[s2]);
    if (mode != 1) for (int r = 0; r < nr; ++r) if (rr[r] != 0) return false;        // receivers on the coarse grid
    if (mode != 2) for (int s2 = 0; s2 < ns; ++s2) if (rs[s2] != 0) return false;    // sources on the coarse grid
    // residues of the fine side -> slots 0..Q-1
    std::vector<long long> resid;
    if (mode == 1) resid.assign(rr.begin(), rr.end());
    else if (mode == 2) resid.assign(rs.begin(), rs.end());
    else resid.push_back(0);
    std::sort(resid.begin(), resid.end());
    resid.erase(std::unique(resid.begin(), resid.end()), resid.end());
    pl.Q = (int)resid.size();
    if (pl.Q < 1 || pl.Q > 6) return false;
    auto slot_of = [&](long long r) { return (int)(std::lower_bound(resid.begin(), resid.end(), r) - resid.begin()); };
    const long long mmin = *std::min_element(mr.begin(), mr.end()), mmax = *std::max_element(mr.begin(), mr.end());
    const long long jmin = *std::min_element(js.begin(), js.end()), jmax = *std::max_element(js.begin(), js.end());
    if (mmax - mmin > (1 << 20) || jmax - jmin > (1 << 20)) return false;
    pl.MR = (int)(mmax - mmin + 1); pl.NS = (int)(jmax - jmin + 1);
    const int npad = npad_for > 0 ? npad_for : pl.NS;
    if (npad < pl.NS) return false;
    // per x class: one (coarse index, slot)
    std::vector<long long> am(g1.nr, LLONG_MIN), bj(g1.ns, LLONG_MIN);
    std::vector<int> aslot(g1.nr, 0), bslot(g1.ns, 0);
    for (int r = 0; r < nr; ++r) {
        const int a = g1.rcls[r]; const int sl = mode == 1 ? slot_of(rr[r]) : 0;
        if (am[a] == LLONG_MIN) { am[a] = mr[r] - mmin; aslot[a] = sl; }
        else if (am[a] != mr[r] - mmin || aslot[a] != sl) return false;
    }
    for (int s2 = 0; s2 < ns; ++s2) {
        const int b = g1.scls[s2]; const int sl = mode == 2 ? slot_of(rs[s2]) : 0;
        if (bj[b] == LLONG_MIN) { bj[b] = js[s2] - jmin; bslot[b] = sl; }
        else if (bj[b] != js[s2] - jmin || bslot[b] != sl) return false;
    }
    const int nd = pl.MR + pl.NS - 1;
    pl.dcls.assign((size_t)6 * nd, -1);
    for (int a = 0; a < g1.nr; ++a) {
        if (am[a] == LLONG_MIN) continue;
        for (int b = 0; b < g1.ns; ++b) {
            if (bj[b] == LLONG_MIN) continue;
            const int slot = mode == 1 ? aslot[a] : mode == 2 ? bslot[b] : 0;
            int& q = pl.dcls[(size_t)slot * nd + (size_t)(am[a] - bj[b] + pl.NS - 1)];
            const int cls = g1.D[(size_t)a * g1.ns + b];
            if (q < 0) q = cls; else if (q != cls) return false;
        }
    }
    // receiver runs
    const int nr23 = pc.g23.nr;
    std::vector<std::vector<int>> groups(nr23);
    for (int r = 0; r < nr; ++r) groups[pc.g23.rcls[r]].push_back(r);
    pl.drow.clear(); pl.dbeg.clear(); pl.dcnt.clear(); pl.dm0.clear(); pl.out_map.clear();
    const int run = blk * kCdG;
    for (int g = 0; g < nr23; ++g) {
        std::vector<int>& mem = groups[g];
        std::sort(mem.begin(), mem.end(), [&](int a, int b) { return mr[a] != mr[b] ? mr[a] < mr[b] : rr[a] < rr[b]; });
        size_t k = 0;
        while (k < mem.size()) {
            // one run: consecutive coarse positions starting at mr[mem[k]]
            const long long m0 = mr[mem[k]];
            const int entry0 = (int)(pl.out_map.size() / 6);
            long long mcur = m0;
            int count = 0;
            while (k < mem.size() && count < run) {
                if (mr[mem[k]] != mcur) break;
                pl.out_map.resize(pl.out_map.size() + 6, -1);
                int* om = &pl.out_map[pl.out_map.size() - 6];
                while (k < mem.size() && mr[mem[k]] == mcur) {                 // the receivers of this coarse position
                    const int r = mem[k];
                    if (mode == 1) {
                        const int sl = slot_of(rr[r]);
                        if (om[sl] >= 0) return false;                         // two receivers, one position
                        om[sl] = r;                                            // K = 1: y index = r
                    } else {
                        if (om[0] >= 0) return false;
                        for (int kk = 0; kk < K; ++kk) om[kk] = kk * nr + r;
                    }
                    ++k;
                }
                ++count; ++mcur;
            }
            pl.drow.push_back(g); pl.dbeg.push_back(entry0); pl.dcnt.push_back(count); pl.dm0.push_back((int)(m0 - mmin));
        }
    }
    if ((long long)pl.drow.size() > 4LL * nr23 + 64) return false;             // positions too scattered to be worth it
    // sources by (group, coarse position, column)
    const int ns23 = pc.g23.ns;
    pl.xmap.assign((size_t)ns23 * npad * 6, -1);
    for (int s2 = 0; s2 < ns; ++s2) {
        int* xm = &pl.xmap[((size_t)pc.g23.scls[s2] * npad + (size_t)(js[s2] - jmin)) * 6];
        if (mode == 2) {
            const int sl = slot_of(rs[s2]);
            if (xm[sl] >= 0) return false;
            xm[sl] = s2;                                                       // P = 1: x index = s
        } else {
            if (xm[0] >= 0) return false;
            for (int c = 0; c < 6; ++c) xm[c] = c * ns + s2;                   // x[p*ns + s]
        }
    }
    return true;
}

// integer positions of coordinates on their common grid: step = smallest gap between distinct values; false when a
// value does not sit on that grid (to 1e-6 of the step) or the grid would be absurdly long
static inline bool grid_positions(const std::vector<double>& x, std::vector<int>& pos)
{
    std::vector<int> idx;
    const int ncl = cluster_values(x, 1e-12 * span_of(x, x), idx);
    if (ncl < 1) return false;
    std::vector<double> val(ncl);
    for (size_t i = 0; i < x.size(); ++i) val[idx[i]] = x[i];
    double step = 0;
    for (int c = 1; c < ncl; ++c) step = step == 0 ? val[c] - val[c - 1] : std::min(step, val[c] - val[c - 1]);
    pos.assign(x.size(), 0);
    if (ncl == 1) return true;
    // the step of the common grid may be a fraction of the smallest gap (e.g. centres of cells 5 and 1 units long):
    // try step / d for small d
    for (int d = 1; d <= 4; ++d) {
        const double h = step / d;
        bool ok = (val[ncl - 1] - val[0]) / h < 4e6;
        std::vector<long long> p(ncl);
        for (int c = 0; c < ncl && ok; ++c) {
            const double t = (val[c] - val[0]) / h;
            p[c] = (long long)std::llround(t);
            if (std::fabs(t - (double)p[c]) > 1e-6) ok = false;
        }
        if (ok) { for (size_t i = 0; i < x.size(); ++i) pos[i] = (int)p[idx[i]]; return true; }
    }
    return false;
}

// origin and step of an arithmetic progression of positions (the coarse grid); false if they are not one
static inline bool coarse_grid(const std::vector<int>& pos, long long& c0, long long& qstep)
{
    std::vector<int> u(pos);
    std::sort(u.begin(), u.end());
    u.erase(std::unique(u.begin(), u.end()), u.end());
    c0 = u[0]; qstep = 1;
    if (u.size() == 1) return true;
    qstep = u[1] - u[0];
    for (size_t i = 1; i < u.size(); ++i) if (u[i] - u[i - 1] != qstep) return false;
    return true;
}

// rpos / spos (optional): integer x positions of the local receivers / all sources on their common grid, with the
// origin c0 and step qstep of the coarse grid (plan_offsets); mode: 0 6x6 one grid, 1 sources coarse, 2 receivers coarse
static inline int make_class_operand(const Hex8PairClasses& pc, const DevBuf<double>& table, int K, int P, int nr, int ns,
                                     ClassOperand& c, const int* rpos = nullptr, const int* spos = nullptr, int mode = 0,
                                     long long c0 = 0, long long qstep = 1, const std::vector<int>* sg_walk = nullptr)
{
    c.K = K; c.P = P; c.nr = nr; c.ns = ns;
    c.n1 = pc.g1.n; c.n23 = pc.g23.n; c.ns1 = pc.g1.ns; c.ns23 = pc.g23.ns;
    const int KP = K * P;
    // 8 consecutive classes x one 128-bit load each must cover the 32 banks exactly once: ts*2 words = 4 (mod 8) words
    // apart -> ts = 6 (6x1, 1x6: 12 words) and 38 (6x6: 76 words) both satisfy it
    c.ts = KP == 36 ? 38 : KP;
    const size_t ncls = (size_t)c.n1 * c.n23;
    OQ_CHECK(table.n >= ncls * KP, "class table is smaller than its maps");
    OQ_TRY(c.Tm.alloc(ncls * c.ts));
    class_table_transpose_kernel<<<(unsigned)((ncls * c.ts + 255) / 256), 256>>>(table.p, ncls, KP, c.ts, c.Tm.p);
    OQ_LAUNCHED();
    OQ_TRY(c.rc1.upload(pc.g1.rcls.data(), pc.g1.rcls.size()));
    OQ_TRY(c.sc1.upload(pc.g1.scls.data(), pc.g1.scls.size()));
    OQ_TRY(c.D1.upload(pc.g1.D.data(), pc.g1.D.size()));
    OQ_TRY(c.D23.upload(pc.g23.D.data(), pc.g23.D.size()));
    OQ_TRY(c.rc23.upload(pc.g23.rcls.data(), pc.g23.rcls.size()));
    OQ_TRY(c.sc23.upload(pc.g23.scls.data(), pc.g23.scls.size()));
    OQ_CHECK((int)pc.g1.rcls.size() == nr && (int)pc.g23.rcls.size() == nr, "class maps do not match the shard's receivers");
    OQ_CHECK((int)pc.g1.scls.size() == ns && (int)pc.g23.scls.size() == ns, "class maps do not match the sources");
    // sources grouped by (y,z) class
    std::vector<int> sptr(c.ns23 + 1, 0), sitems(ns);
    for (int s = 0; s < ns; ++s) ++sptr[pc.g23.scls[s] + 1];
    for (int g = 0; g < c.ns23; ++g) sptr[g + 1] += sptr[g];
    {
        std::vector<int> fill(sptr.begin(), sptr.end() - 1);
        for (int s = 0; s < ns; ++s) sitems[fill[pc.g23.scls[s]]++] = s;
    }
    c.max_sg = 0;
    for (int g = 0; g < c.ns23; ++g) c.max_sg = std::max(c.max_sg, sptr[g + 1] - sptr[g]);
    // receivers ordered by (y,z) class; one CTA per run of <= rb receivers of a class
    const int nr23 = pc.g23.nr;
    std::vector<int> rptr(nr23 + 1, 0), ritems(nr > 0 ? nr : 1);
    for (int r = 0; r < nr; ++r) ++rptr[pc.g23.rcls[r] + 1];
    for (int g = 0; g < nr23; ++g) rptr[g + 1] += rptr[g];
    {
        std::vector<int> fill(rptr.begin(), rptr.end() - 1);
        for (int r = 0; r < nr; ++r) ritems[fill[pc.g23.rcls[r]]++] = r;
    }
    // within a (y,z) class order the receivers by the x class they form with the first source: consecutive lanes then
    // read consecutive class blocks (on commensurate grids receivers of one residue sit together), which keeps the
    // 128-bit loads of a warp spread over the banks.  The order of the receivers inside a CTA changes no sum.
    for (int g = 0; g < nr23; ++g)
        std::stable_sort(ritems.begin() + rptr[g], ritems.begin() + rptr[g + 1], [&](int a, int b) {
            return pc.g1.D[(size_t)pc.g1.rcls[a] * pc.g1.ns] < pc.g1.D[(size_t)pc.g1.rcls[b] * pc.g1.ns];
        });
    int maxcount = 0;
    for (int g = 0; g < nr23; ++g) maxcount = std::max(maxcount, rptr[g + 1] - rptr[g]);
    auto ctas_for = [&](int rb) { long long n = 0; for (int g = 0; g < nr23; ++g) n += (rptr[g + 1] - rptr[g] + rb - 1) / rb; return n; };
    // a fixed run length: the number of source slices (256 / rb) fixes the association order of a receiver's sum, which
    // must not depend on the shard
    const int rb = 64;                                          // = kCmRb of classmat.cuh
    (void)maxcount; (void)ctas_for;
    std::vector<int> crow, cbeg, ccnt;
    for (int g = 0; g < nr23; ++g)
        for (int b = rptr[g]; b < rptr[g + 1]; b += rb) {
            crow.push_back(g); cbeg.push_back(b); ccnt.push_back(std::min(rb, rptr[g + 1] - b));
        }
    c.nctas = (int)crow.size();
    // Every row of D23 walks the source groups in their natural order (the (y,z) classes of the sources, numbered as the
    // mesh lists them).  On a structured mesh the slab a receiver row (ry, rz) needs for the source row (sy, sz) is the
    // one of the offset ry - sy: the neighbouring receiver row needs it ONE step later, and since every row has the
    // same number of sources per layer the lag never grows -- rows that run at the same time share their fetches through
    // L2.  (Ordering by class id instead let the rows drift apart by hundreds of steps: L2 hit rate 6 %.)
    {
        // sg_walk (optional, a permutation of the source groups that depends on the mesh alone): the order of the walk
        // -- for the mantle sources layer by layer, along y inside a layer, so that the lag between neighbouring receiver
        // rows is one step, not one sweep over the layers
        std::vector<int> order((size_t)nr23 * c.ns23);
        const bool walk = sg_walk && (int)sg_walk->size() == c.ns23;
        for (int g = 0; g < nr23; ++g)
            for (int b = 0; b < c.ns23; ++b) order[(size_t)g * c.ns23 + b] = walk ? (*sg_walk)[b] : b;
        OQ_TRY(c.sg_order.upload(order.data(), order.size()));
    }
    OQ_TRY(c.rg_items.upload(ritems.data(), ritems.size()));
    OQ_TRY(c.sg_ptr.upload(sptr.data(), sptr.size()));
    if (c.nctas) {
        OQ_TRY(c.cta_row.upload(crow.data(), crow.size()));
        OQ_TRY(c.cta_begin.upload(cbeg.data(), cbeg.size()));
        OQ_TRY(c.cta_count.upload(ccnt.data(), ccnt.size()));
    }
    const int PX = (P + 1) & ~1;
    // the sources of every group in slot order (padded to a multiple of 8 slots): what class_gather_x_kernel gathers by
    // and the x classes the CTAs fetch beside the forcing values
    c.xstride = (int)round_up((size_t)std::max(c.max_sg, 1), 8);
    {
        std::vector<int> xmap((size_t)c.ns23 * c.xstride, -1), csg((size_t)c.ns23 * c.xstride, 0);
        for (int g = 0; g < c.ns23; ++g)
            for (int j = sptr[g]; j < sptr[g + 1]; ++j) {
                xmap[(size_t)g * c.xstride + (j - sptr[g])] = sitems[j];
                csg[(size_t)g * c.xstride + (j - sptr[g])] = pc.g1.scls[sitems[j]];
            }
        OQ_TRY(c.xmap.upload(xmap.data(), xmap.size()));
        OQ_TRY(c.csg.upload(csg.data(), csg.size()));
        OQ_TRY(c.xg.alloc((size_t)c.ns23 * c.xstride * PX));
        OQ_TRY(c.part.alloc((size_t)4 * K * (nr > 0 ? nr : 1)));
    }
    const size_t stage = (size_t)c.n1 * c.ts * sizeof(double) + (size_t)c.xstride * PX * sizeof(double) + (size_t)c.xstride * sizeof(int);
    c.smem = 2 * stage + (size_t)256 * K * sizeof(double);
    {   // the receivers' rows of D1 (64 rows, odd stride) behind the stages, if they fit and leave room for 2 CTAs per SM
        const size_t d1b = round_up((size_t)64 * cm_d1_stride(c.ns1) * sizeof(unsigned short), 16);
        c.d1_smem = c.n1 < 65536 && c.smem + d1b <= 110 * 1024;
        if (c.d1_smem) c.smem += d1b;
    }
    OQ_CHECK(c.smem <= 226 * 1024, "class form: %d x-classes of %d doubles do not fit shared memory (%zu bytes)", c.n1, c.ts, c.smem);
    c.table_bytes = (double)ncls * c.ts * sizeof(double);
    if (rpos && spos && nr > 0) {
        // the sliding-window kernel, if the x classes are a function of (residue, coarse offset)
        OffsetPlan pl;
        bool ok = plan_offsets(pc, mode, rpos, spos, nr, ns, c0, qstep, K, kCdBlk, 0, pl);      // first pass: NS, MR
        if (ok) {
            c.dL = (int)round_up((size_t)(pl.NS + kCdSlices - 1) / kCdSlices, kCdG);
            const int npad = kCdSlices * c.dL;
            c.dblk = kCdBlk;
            ok = plan_offsets(pc, mode, rpos, spos, nr, ns, c0, qstep, K, c.dblk, npad, pl);
            if (ok && pl.drow.size() * 4 < 3 * 148) {     // too few runs of 64 (x 4 source quarters) to fill the GPU: runs of 32
                c.dblk = 4;
                ok = plan_offsets(pc, mode, rpos, spos, nr, ns, c0, qstep, K, c.dblk, npad, pl);
            }
            const size_t ndp = (size_t)c.dblk * kCdG + npad, ndp8 = (size_t)kCdBlk * kCdG + npad;
            c.noff = (int)(pl.MR - 1 + ndp8);
            c.dsmem = 2 * (ndp * 38 + ndp / 2 + (size_t)npad * 6) * sizeof(double) + (size_t)kCdSlices * c.dblk * kCdG * 6 * sizeof(double);
            size_t free_b = 0, total_b = 0;
            const double td_bytes = (double)c.n23 * c.noff * 38 * sizeof(double);
            const bool fits = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && td_bytes < 0.5 * (double)free_b;
            if (ok && fits && c.dsmem <= 226 * 1024) {
                const int nd = pl.MR + pl.NS - 1;
                DevBuf<int> ddcls;
                OQ_TRY(ddcls.upload(pl.dcls.data(), pl.dcls.size()));
                const size_t tot = (size_t)c.n23 * c.noff * 38;
                OQ_TRY(c.Td.alloc(tot));
                class_table_offset_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(c.Tm.p, c.ts, ddcls.p, nd, mode, pl.Q, (size_t)c.n23, c.n1,
                                                                                  c.noff, npad - pl.NS, c.Td.p);
                OQ_LAUNCHED();
                OQ_CUDA(cudaDeviceSynchronize());
                OQ_TRY(c.dxmap.upload(pl.xmap.data(), pl.xmap.size()));
                OQ_TRY(c.dxg.alloc((size_t)c.ns23 * npad * 6));
                OQ_TRY(c.dout_map.upload(pl.out_map.data(), pl.out_map.size()));
                OQ_TRY(c.dcta_row.upload(pl.drow.data(), pl.drow.size()));
                OQ_TRY(c.dcta_begin.upload(pl.dbeg.data(), pl.dbeg.size()));
                OQ_TRY(c.dcta_count.upload(pl.dcnt.data(), pl.dcnt.size()));
                OQ_TRY(c.dcta_m0.upload(pl.dm0.data(), pl.dm0.size()));
                c.ndctas = (int)pl.drow.size();
                c.npos = pl.NS; c.dmode = mode; c.dQ = pl.Q;
                c.diag_ok = true;
                c.table_bytes += td_bytes;
            }
        }
    }
    OQ_CUDA(cudaDeviceSynchronize());
    return 0;
}

// dense rows of a class-form shard (parity checks, oq_matrix_to_host): G [K*nr x ld] row-major
static inline int expand_class_operand(const ClassOperand& c, size_t ld, double* G)
{
    ClassView v{};
    v.rc1 = c.rc1.p; v.rc23 = c.rc23.p; v.sc1 = c.sc1.p; v.sc23 = c.sc23.p; v.D1 = c.D1.p; v.D23 = c.D23.p;
    v.ns1 = c.ns1; v.ns23 = c.ns23; v.n1 = c.n1; v.n23 = c.n23;
    if (c.nr == 0) return 0;
    dim3 grid((unsigned)std::min<size_t>(((size_t)c.ns + 255) / 256, 64), (unsigned)std::min(c.nr, 65535));
    if (c.K == 6 && c.P == 6) expand_classes_kernel<6, 6><<<grid, 256>>>(c.Tm.p, v, c.nr, c.ns, ld, G, (size_t)c.ts, 1);
    else if (c.K == 6 && c.P == 1) expand_classes_kernel<6, 1><<<grid, 256>>>(c.Tm.p, v, c.nr, c.ns, ld, G, (size_t)c.ts, 1);
    else if (c.K == 1 && c.P == 6) expand_classes_kernel<1, 6><<<grid, 256>>>(c.Tm.p, v, c.nr, c.ns, ld, G, (size_t)c.ts, 1);
    else return fail("class-form operand with %dx%d blocks is not supported", c.K, c.P);
    OQ_LAUNCHED();
    return 0;
}

}  // namespace oq
