// amge_topology.hpp -- host-side integer structures of the coarsening path:
//   signed incidence tables (TopologyTable)          src/topology/TopologyTable.cpp:97-139
//   AgglomeratedTopology::CoarsenLocalPartitioning   src/topology/Topology.cpp:685-828
//   findMinimalIntersectionSets                      src/structures/minimalIntersectionSet.cpp:43-130
//   MFEMRefinedMeshPartitioner::Partition            src/partitioning/MFEMRefinedMeshPartitioner.cpp:48-90
//   connectedComponents                              src/structures/connectedComponents.cpp:23-87
//   AgglomeratedTopologyCheck                        src/topology/AgglomeratedTopologyCheck.cpp:25-316
//   DeAgglomerateBadAgglomeratedEntities             src/topology/Topology.cpp:1151-1214
// These are O(n) integer graph operations that run once per level (bit-exact
// required; host in this round).  All tables are kept in canonical CSR form (ascending
// column indices per row) -- the numbering convention shared with oracle/amge.py.
#pragma once
#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>
#include "parelag_sequence.hpp"

namespace parelag
{
namespace hostcsr
{
/// C = A*B, rows sorted ascending, exact cancellations (|c| < tol) dropped when drop_tol >= 0.
/// Rows are independent: every thread owns a contiguous row range with its own marker/accumulator,
/// row products land in per-thread buffers and are stitched together in row order, so the result
/// (pattern, value bits) does not depend on the thread count.
inline HostCSR Mult(const HostCSR &A, const HostCSR &B, double drop_tol = -1.0)
{
    PARELAG_TEST_FOR_EXCEPTION(A.ncols != B.nrows, std::logic_error, "hostcsr::Mult: size mismatch");
    HostCSR C;
    C.nrows = A.nrows; C.ncols = B.ncols;
    C.I.assign(A.nrows + 1, 0);
    int nt = 1;
#ifdef _OPENMP
    nt = std::max(1, std::min(omp_get_max_threads(), A.nrows / 4096 + 1));
#endif
    std::vector<std::vector<int>> Jt(nt);
    std::vector<std::vector<double>> At(nt);
    std::vector<int> r0(nt + 1, 0);
    for (int t = 0; t <= nt; ++t) r0[t] = (int)((int64_t)A.nrows * t / nt);
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        std::vector<int> marker(B.ncols, -1), cols;
        std::vector<double> acc(B.ncols, 0.0);
        std::vector<int> &Jo = Jt[t];
        std::vector<double> &Ao = At[t];
        for (int i = r0[t]; i < r0[t + 1]; ++i)
        {
            cols.clear();
            for (int ka = A.I[i]; ka < A.I[i + 1]; ++ka)
            {
                const int k = A.J[ka];
                const double a = A.A[ka];
                for (int kb = B.I[k]; kb < B.I[k + 1]; ++kb)
                {
                    const int j = B.J[kb];
                    if (marker[j] != i) { marker[j] = i; cols.push_back(j); acc[j] = a * B.A[kb]; }
                    else acc[j] += a * B.A[kb];
                }
            }
            std::sort(cols.begin(), cols.end());
            int cnt = 0;
            for (int j : cols)
                if (drop_tol < 0.0 || std::fabs(acc[j]) >= drop_tol) { Jo.push_back(j); Ao.push_back(acc[j]); ++cnt; }
            C.I[i + 1] = cnt;
        }
    }
    for (int i = 0; i < A.nrows; ++i) C.I[i + 1] += C.I[i];
    C.J.resize((size_t)C.I[A.nrows]); C.A.resize((size_t)C.I[A.nrows]);
#pragma omp parallel for num_threads(nt) schedule(static, 1)
    for (int t = 0; t < nt; ++t)
    {
        std::copy(Jt[t].begin(), Jt[t].end(), C.J.begin() + C.I[r0[t]]);
        std::copy(At[t].begin(), At[t].end(), C.A.begin() + C.I[r0[t]]);
    }
    return C;
}
/// boolean product: pattern of A*B with unit values, rows sorted ascending (same threading scheme as Mult)
inline HostCSR MultPattern(const HostCSR &A, const HostCSR &B)
{
    PARELAG_TEST_FOR_EXCEPTION(A.ncols != B.nrows, std::logic_error, "hostcsr::MultPattern: size mismatch");
    HostCSR C;
    C.nrows = A.nrows; C.ncols = B.ncols;
    C.I.assign(A.nrows + 1, 0);
    int nt = 1;
#ifdef _OPENMP
    nt = std::max(1, std::min(omp_get_max_threads(), A.nrows / 4096 + 1));
#endif
    std::vector<std::vector<int>> Jt(nt);
    std::vector<int> r0(nt + 1, 0);
    for (int t = 0; t <= nt; ++t) r0[t] = (int)((int64_t)A.nrows * t / nt);
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        std::vector<int> marker(B.ncols, -1);
        std::vector<int> &Jo = Jt[t];
        for (int i = r0[t]; i < r0[t + 1]; ++i)
        {
            const size_t first = Jo.size();
            for (int ka = A.I[i]; ka < A.I[i + 1]; ++ka)
            {
                const int k = A.J[ka];
                for (int kb = B.I[k]; kb < B.I[k + 1]; ++kb)
                {
                    const int j = B.J[kb];
                    if (marker[j] != i) { marker[j] = i; Jo.push_back(j); }
                }
            }
            std::sort(Jo.begin() + first, Jo.end());
            C.I[i + 1] = (int)(Jo.size() - first);
        }
    }
    for (int i = 0; i < A.nrows; ++i) C.I[i + 1] += C.I[i];
    C.J.resize((size_t)C.I[A.nrows]);
    C.A.assign((size_t)C.I[A.nrows], 1.0);
#pragma omp parallel for num_threads(nt) schedule(static, 1)
    for (int t = 0; t < nt; ++t) std::copy(Jt[t].begin(), Jt[t].end(), C.J.begin() + C.I[r0[t]]);
    return C;
}
inline HostCSR Transpose(const HostCSR &A)
{
    HostCSR T;
    T.nrows = A.ncols; T.ncols = A.nrows;
    T.I.assign(A.ncols + 1, 0);
    T.J.resize(A.J.size()); T.A.resize(A.A.size());
    for (int c : A.J) T.I[c + 1]++;
    for (int c = 0; c < A.ncols; ++c) T.I[c + 1] += T.I[c];
    std::vector<int> next(T.I.begin(), T.I.end() - 1);
    for (int i = 0; i < A.nrows; ++i)
        for (int k = A.I[i]; k < A.I[i + 1]; ++k) { int p = next[A.J[k]]++; T.J[p] = i; T.A[p] = A.A[k]; }
    return T;
}
inline HostCSR Add(const HostCSR &A, const HostCSR &B)
{
    HostCSR C;
    C.nrows = A.nrows; C.ncols = A.ncols;
    C.I.assign(A.nrows + 1, 0);
    for (int i = 0; i < A.nrows; ++i)
    {
        int ka = A.I[i], kb = B.I[i];
        while (ka < A.I[i + 1] || kb < B.I[i + 1])
        {
            int ja = ka < A.I[i + 1] ? A.J[ka] : INT32_MAX, jb = kb < B.I[i + 1] ? B.J[kb] : INT32_MAX;
            if (ja == jb) { C.J.push_back(ja); C.A.push_back(A.A[ka++] + B.A[kb++]); }
            else if (ja < jb) { C.J.push_back(ja); C.A.push_back(A.A[ka++]); }
            else { C.J.push_back(jb); C.A.push_back(B.A[kb++]); }
        }
        C.I[i + 1] = (int)C.J.size();
    }
    return C;
}
inline HostCSR Abs(HostCSR A) { for (auto &v : A.A) v = std::fabs(v); return A; }
inline HostCSR Sign(HostCSR A) { for (auto &v : A.A) v = v > 0 ? 1.0 : -1.0; return A; }
inline HostCSR Ones(HostCSR A) { for (auto &v : A.A) v = 1.0; return A; }
/// MultOrientation: product, drop |.| < 1e-10, keep the sign
inline HostCSR MultOrientation(const HostCSR &A, const HostCSR &B) { return Sign(Mult(A, B, 1e-10)); }
inline HostCSR Identity(int n)
{
    HostCSR E;
    E.nrows = E.ncols = n;
    E.I.resize(n + 1); E.J.resize(n); E.A.assign(n, 1.0);
    std::iota(E.I.begin(), E.I.end(), 0);
    std::iota(E.J.begin(), E.J.end(), 0);
    return E;
}
} // namespace hostcsr

/// entity x MIS table (+-1) : entities that belong to exactly the same set of
/// agglomerates, with the same relative orientation, form one coarse entity
inline HostCSR findMinimalIntersectionSets(const HostCSR &Z, double skipDiagEntryLessThan)
{
    const double tol = 1e-10;
    const int n = Z.nrows;
    std::vector<double> diag(n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int k = Z.I[i]; k < Z.I[i + 1]; ++k) if (Z.J[k] == i) diag[i] = Z.A[k];
    std::vector<int> mis_of(n, -1);
    std::vector<double> sign_of(n, 0.0);
    int current = 0;
    for (int i = 0; i < n; ++i)
    {
        const double Zii = diag[i];
        if (Zii - skipDiagEntryLessThan > -tol && mis_of[i] == -1)
        {
            for (int k = Z.I[i]; k < Z.I[i + 1]; ++k)
            {
                const int j = Z.J[k];
                const double Zij = Z.A[k];
                if (std::fabs(diag[j] - Zii) < tol && (std::fabs(Zij - Zii) < tol || std::fabs(Zij + Zii) < tol))
                { mis_of[j] = current; sign_of[j] = Zij / Zii; }
            }
            ++current;
        }
    }
    HostCSR E;
    E.nrows = n; E.ncols = current;
    E.I.assign(n + 1, 0);
    for (int i = 0; i < n; ++i)
    {
        if (mis_of[i] >= 0) { E.J.push_back(mis_of[i]); E.A.push_back(sign_of[i]); }
        E.I[i + 1] = (int)E.J.size();
    }
    return E;
}

/// The same minimal intersection sets computed from membership signatures, in O(nnz):
/// with Z = X^T X for a 0/+-1 table X (agglomerate x entity), entities i and j satisfy the
/// MIS criterion (Z_ii == Z_jj and |Z_ij| == Z_ii) iff they belong to exactly the same
/// agglomerates with one consistent relative orientation.  `memb` is X^T (entity x
/// agglomerate, ascending columns).  Avoids forming Z, whose boundary-attribute part
/// facet_bdr * facet_bdr^T grows quadratically (Topology.cpp:748-757 warns about it; at
/// 144^3 hexahedra it has 2.6e9 entries).  MIS are numbered by their first entity, exactly
/// like findMinimalIntersectionSets.
inline HostCSR MinimalIntersectionSetsFromMembership(const HostCSR &memb)
{
    const int n = memb.nrows;
    std::vector<int> mis_of(n, -1);
    std::vector<double> sign_of(n, 0.0);
    // open-addressing table of MIS ids keyed by the (agglomerate, relative sign) signature; a slot is verified against
    // the signature of the set's first entity, so hash collisions cannot merge sets
    size_t cap = 16;
    while (cap < (size_t)n * 2) cap <<= 1;
    std::vector<int> table(cap, -1), first_entity;
    first_entity.reserve((size_t)n / 4 + 16);
    auto relsign = [&](int k, double s0) { return ((memb.A[k] > 0 ? 1.0 : -1.0) * s0 > 0) ? 1u : 0u; };
    int current = 0;
    for (int i = 0; i < n; ++i)
    {
        const int lo = memb.I[i], hi = memb.I[i + 1];
        if (hi == lo) continue;                       // diag(Z) < 0.5: belongs to no MIS
        const double s0 = memb.A[lo] > 0 ? 1.0 : -1.0;
        uint64_t h = 0x9e3779b97f4a7c15ull ^ (uint64_t)(hi - lo);
        for (int k = lo; k < hi; ++k)
        {
            h ^= ((uint64_t)(uint32_t)memb.J[k] << 1) | relsign(k, s0);
            h *= 0xff51afd7ed558ccdull;
            h ^= h >> 33;
        }
        size_t slot = (size_t)h & (cap - 1);
        int id = -1;
        for (;; slot = (slot + 1) & (cap - 1))
        {
            const int cand = table[slot];
            if (cand < 0) break;
            const int f = first_entity[cand], flo = memb.I[f], fhi = memb.I[f + 1];
            if (fhi - flo != hi - lo) continue;
            const double f0 = memb.A[flo] > 0 ? 1.0 : -1.0;
            bool same = true;
            for (int k = 0; k < hi - lo && same; ++k)
                same = memb.J[lo + k] == memb.J[flo + k] && relsign(lo + k, s0) == relsign(flo + k, f0);
            if (same) { id = cand; break; }
        }
        if (id < 0) { id = current++; table[slot] = id; first_entity.push_back(i); }
        mis_of[i] = id;
        sign_of[i] = s0;      // orientation relative to the first entity of the set (whose s0-normalised pattern is all '+' ... )
    }
    // orientation of entity i relative to the first entity f of its set: Z_fi / Z_ff = s0(i) * s0(f)
    std::vector<double> first_sign(current, 0.0);
    for (int i = 0; i < n; ++i)
        if (mis_of[i] >= 0 && first_sign[mis_of[i]] == 0.0) first_sign[mis_of[i]] = sign_of[i];
    HostCSR E;
    E.nrows = n; E.ncols = current;
    E.I.assign(n + 1, 0);
    for (int i = 0; i < n; ++i)
    {
        if (mis_of[i] >= 0) { E.J.push_back(mis_of[i]); E.A.push_back(sign_of[i] * first_sign[mis_of[i]]); }
        E.I[i + 1] = (int)E.J.size();
    }
    return E;
}

/// connectedComponents (src/structures/connectedComponents.cpp:23-87): every partition is split into its connected
/// components with respect to the element-element table (two elements are neighbours when they share a facet:
/// LocalElementElementTable = B_0 W B_0^T, Topology.cpp:281-293; here the facets are walked directly).  Component c of
/// partition p becomes offset[p] + c, the components of a partition numbered in the order a scan of the elements meets
/// them: empty partitions disappear, connected partitions keep their relative order.  Returns the number of agglomerates.
inline int ConnectedComponents(std::vector<int> &partitioning, const HostCSR &el_facet, const HostCSR &facet_el)
{
    const int n = (int)partitioning.size();
    if (n == 0) return 0;
    int npart = 0;
    for (int p : partitioning) npart = std::max(npart, p + 1);
    std::vector<int> component((size_t)n, -1), offset((size_t)npart + 1, 0), stack;
    stack.reserve(1024);
    for (int node = 0; node < n; ++node)
    {
        if (partitioning[node] < 0 || component[node] >= 0) continue;
        component[node] = offset[partitioning[node] + 1]++;
        stack.assign(1, node);
        while (!stack.empty())
        {
            const int i = stack.back();
            stack.pop_back();
            for (int kf = el_facet.I[i]; kf < el_facet.I[i + 1]; ++kf)
            {
                const int f = el_facet.J[kf];
                for (int ke = facet_el.I[f]; ke < facet_el.I[f + 1]; ++ke)
                {
                    const int k = facet_el.J[ke];
                    if (partitioning[k] == partitioning[i] && component[k] < 0) { component[k] = component[i]; stack.push_back(k); }
                }
            }
        }
    }
    for (int p = 0; p < npart; ++p) offset[p + 1] += offset[p];
    for (int i = 0; i < n; ++i)
        if (partitioning[i] >= 0) partitioning[i] = offset[partitioning[i]] + component[i];
    return offset[npart];
}

/// AgglomeratedTopologyCheck (src/topology/AgglomeratedTopologyCheck.cpp): Betti numbers of the agglomerated entities,
/// the boundary-connectivity check, the messages the reference prints and the de-agglomeration of bad entities.
/// The functions take the fine boundary operators B and the agglomerated-entity tables built so far
/// (AEntity_entity[0..codim]) instead of the topology object, so that CoarsenLocalPartitioning can call them stage by stage.
struct AgglomeratedTopologyCheck
{
    /// numerical rank of a dense m x n matrix (row major, destroyed): Gaussian elimination with complete pivoting, pivots
    /// below tol count as zero (mfem::DenseMatrix::Rank(1e-9) in the reference; the matrices are signed incidence matrices)
    static int DenseRank(std::vector<double> &a, int m, int n, double tol)
    {
        int rank = 0;
        std::vector<int> rows((size_t)m), cols((size_t)n);
        std::iota(rows.begin(), rows.end(), 0);
        std::iota(cols.begin(), cols.end(), 0);
        for (; rank < std::min(m, n); ++rank)
        {
            int pi = -1, pj = -1;
            double best = tol;
            for (int i = rank; i < m; ++i)
                for (int j = rank; j < n; ++j)
                {
                    const double v = std::fabs(a[(size_t)rows[i] * n + cols[j]]);
                    if (v > best) { best = v; pi = i; pj = j; }
                }
            if (pi < 0) break;
            std::swap(rows[rank], rows[pi]);
            std::swap(cols[rank], cols[pj]);
            const double *pr = &a[(size_t)rows[rank] * n];
            const double piv = pr[cols[rank]];
            for (int i = rank + 1; i < m; ++i)
            {
                double *ri = &a[(size_t)rows[i] * n];
                const double f = ri[cols[rank]] / piv;
                if (f == 0.0) continue;
                for (int j = rank; j < n; ++j) ri[cols[j]] -= f * pr[cols[j]];
            }
        }
        return rank;
    }
    /// computeBettiNumbersAgglomeratedEntities (:242-316): betti[a * nlow + l], l = 0 (connected components) .. nlow - 1
    static std::vector<int> ComputeBettiNumbers(int ndim, const std::vector<HostCSR> &B, const std::vector<HostCSR> &AEe, int codim, int &nlow)
    {
        nlow = ndim - codim;
        if (nlow <= 0) { nlow = 0; return {}; }
        std::vector<HostCSR> tabs;
        tabs.push_back(AEe[codim]);
        for (int i = 0; i < nlow; ++i) tabs.push_back(hostcsr::MultPattern(tabs[i], B[codim + i]));
        const int nAE = tabs[0].nrows;
        std::vector<int> betti((size_t)nAE * nlow, 0), dim_k((size_t)nlow + 1), rank_k((size_t)nlow + 1);
        std::vector<double> dloc;
        std::vector<int> colpos;
        for (int a = 0; a < nAE; ++a)
        {
            for (int i = 0; i <= nlow; ++i) dim_k[i] = tabs[i].I[a + 1] - tabs[i].I[a];
            rank_k[nlow] = 0;
            for (int i = 0; i < nlow; ++i)
            {
                rank_k[i] = 0;
                if (dim_k[i] == 0 || dim_k[i + 1] == 0) continue;
                const HostCSR &Bi = B[codim + i];
                const int *rr = &tabs[i].J[tabs[i].I[a]], *cc = &tabs[i + 1].J[tabs[i + 1].I[a]];
                dloc.assign((size_t)dim_k[i] * dim_k[i + 1], 0.0);
                for (int r = 0; r < dim_k[i]; ++r)
                    for (int k = Bi.I[rr[r]]; k < Bi.I[rr[r] + 1]; ++k)
                    {
                        const int *pos = std::lower_bound(cc, cc + dim_k[i + 1], Bi.J[k]);
                        if (pos != cc + dim_k[i + 1] && *pos == Bi.J[k]) dloc[(size_t)r * dim_k[i + 1] + (pos - cc)] = Bi.A[k];
                    }
                rank_k[i] = DenseRank(dloc, dim_k[i], dim_k[i + 1], 1e-9);
            }
            for (int i = 0; i < nlow; ++i) betti[(size_t)a * nlow + (nlow - i - 1)] = dim_k[i + 1] - rank_k[i] - rank_k[i + 1];
        }
        return betti;
    }
    /// additionalTopologyCheck (:25-82): on the boundary of an agglomerated element (codim 0) / facet (codim 1) every
    /// boundary ridge (peak) must be adjacent to exactly two boundary facets (ridges)
    static void AdditionalTopologyCheck(const std::vector<HostCSR> &B, const std::vector<HostCSR> &AEe, int codim, std::vector<int> &isbad,
                                        std::vector<std::string> *messages)
    {
        // interior entities cancel in the signed product; exact zeros are dropped
        const HostCSR bf = hostcsr::Mult(AEe[codim], B[codim], 0.5);
        const HostCSR &fe = B[codim + 1];
        const HostCSR be = hostcsr::MultPattern(bf, fe);
        std::vector<int> count((size_t)fe.ncols, 0);
        for (int a = 0; a < bf.nrows; ++a)
        {
            // twos = (number of boundary facets adjacent to each boundary ridge); the reference tests their sum only
            long sum = 0;
            for (int k = bf.I[a]; k < bf.I[a + 1]; ++k) sum += fe.I[bf.J[k] + 1] - fe.I[bf.J[k]];
            const long nridges = be.I[a + 1] - be.I[a];
            if (sum != 2 * nridges)
            {
                if (messages)
                    messages->push_back("    codim " + std::to_string(codim) + " iAE " + std::to_string(a) +
                                        " has bad connectivity (eg boundary edge adjacent to >2 boundary faces).");
                isbad[a] = 1;
            }
        }
    }
    static bool HasAdditionalCheck(int ndim, int codim) { return (ndim == 2 && codim == 0) || (ndim == 3 && (codim == 0 || codim == 1)); }
    /// MarkBadAgglomeratedEntities (:84-142)
    static bool MarkBadAgglomeratedEntities(int ndim, const std::vector<HostCSR> &B, const std::vector<HostCSR> &AEe, int codim, std::vector<int> &isbad)
    {
        int nlow;
        const std::vector<int> betti = ComputeBettiNumbers(ndim, B, AEe, codim, nlow);
        const int nAE = nlow ? (int)(betti.size() / nlow) : 0;
        isbad.assign((size_t)nAE, 0);
        if (codim <= 2)
            for (int a = 0; a < nAE; ++a)
            {
                if (betti[(size_t)a * nlow] != 1) isbad[a] = 1;                                        // disconnected
                if (codim <= 1)
                    for (int i = 1; i < nlow; ++i) if (betti[(size_t)a * nlow + i] != 0) isbad[a] = 1;  // hole / tunnel
            }
        if (HasAdditionalCheck(ndim, codim)) AdditionalTopologyCheck(B, AEe, codim, isbad, nullptr);
        return std::accumulate(isbad.begin(), isbad.end(), 0) > 0;
    }
    /// ShowBadAgglomeratedEntities and showBadAgglomerated{Elements,Facets,Ridges} (:144-240): the reference's lines
    static void ShowBadAgglomeratedEntities(int ndim, const std::vector<HostCSR> &B, const std::vector<HostCSR> &AEe, int codim,
                                            std::vector<std::string> &out)
    {
        int nlow;
        const std::vector<int> betti = ComputeBettiNumbers(ndim, B, AEe, codim, nlow);
        const int nAE = nlow ? (int)(betti.size() / nlow) : 0;
        static const char *name[3] = {"Element", "Facet", "Ridge"};
        if (codim <= 2)
            for (int a = 0; a < nAE; ++a)
            {
                if (betti[(size_t)a * nlow] != 1)
                    out.push_back(std::string("    ") + name[codim] + " " + std::to_string(a) + " is disconnected. The number of connected components is " +
                                  std::to_string(betti[(size_t)a * nlow]));
                if (codim <= 1)
                    for (int i = 1; i < nlow; ++i)
                        if (betti[(size_t)a * nlow + i] != 0)
                            out.push_back(std::string("    ") + name[codim] + " " + std::to_string(a) + " has " + std::to_string(betti[(size_t)a * nlow + i]) +
                                          ((codim == 1 || i == ndim - 1) ? " holes." : " tunnels."));
            }
        if (HasAdditionalCheck(ndim, codim))
        {
            std::vector<int> dummy((size_t)nAE, 0);
            AdditionalTopologyCheck(B, AEe, codim, dummy, &out);
        }
    }
    /// DeAgglomerateBadAgglomeratedEntities (Topology.cpp:1151-1214): every fine entity of a bad agglomerated entity becomes
    /// an agglomerated entity of its own, in place (the later ones are renumbered)
    static HostCSR DeAgglomerate(const HostCSR &AEE, const std::vector<int> &isbad)
    {
        HostCSR N;
        N.ncols = AEE.ncols; N.J = AEE.J; N.A = AEE.A;
        N.I.assign(1, 0);
        for (int a = 0; a < AEE.nrows; ++a)
        {
            if (isbad[a]) for (int k = AEE.I[a] + 1; k <= AEE.I[a + 1]; ++k) N.I.push_back(k);
            else N.I.push_back(AEE.I[a + 1]);
        }
        N.nrows = (int)N.I.size() - 1;
        return N;
    }
};

class AgglomeratedTopology : public std::enable_shared_from_this<AgglomeratedTopology>
{
public:
    enum Entity { ELEMENT = 0, FACET = 1, RIDGE = 2, PEAK = 3 };
    AgglomeratedTopology(std::vector<HostCSR> B, HostCSR facet_bdr, int ndim)
        : nDim_(ndim), B_(std::move(B)), facet_bdrAttribute_(std::move(facet_bdr))
    {
        n_.push_back(B_[0].nrows);
        for (auto &b : B_) n_.push_back(b.ncols);
    }
    int Dimensions() const { return nDim_; }
    int GetNumberLocalEntities(int codim) const { return n_.at(codim); }
    const HostCSR &GetB(int codim) const { return B_.at(codim); }
    const HostCSR &FacetBdrAttribute() const { return facet_bdrAttribute_; }
    bool HasBdrAttributes() const { return facet_bdrAttribute_.nrows > 0; }
    const HostCSR &AEntityEntity(int codim) const { return AEntity_entity_.at(codim); }
    const std::vector<int> &Partitioning() const { return Partition_; }
    std::shared_ptr<AgglomeratedTopology> CoarserTopology() const { return CoarserTopology_.lock(); }
    std::shared_ptr<AgglomeratedTopology> FinerTopology() const { return FinerTopology_.lock(); }

    /// boolean entity -> sub-entity table |B_big| ... |B_{small-1}| (BuildConnectivity)
    const HostCSR &GetConnectivity(int big, int small) const
    {
        auto key = std::make_pair(big, small);
        auto it = conn_.find(key);
        if (it != conn_.end()) return it->second;
        // |B_big| ... |B_{small-1}| has no cancellations, so the table is the boolean product of the patterns;
        // the chain reuses the cached (big, small-1) table
        if (small == big + 1) return conn_.emplace(key, hostcsr::Ones(B_[big])).first->second;
        const HostCSR &head = GetConnectivity(big, small - 1);
        return conn_.emplace(key, hostcsr::MultPattern(head, B_[small - 1])).first->second;
    }

    /// messages of the last CoarsenLocalPartitioning(..., check_topology = true), one entry per line the reference prints
    const std::vector<std::string> &Messages() const { return messages_; }
    /// AgglomeratedTopology::ShowMe (Topology.cpp:310-352), one rank: entity counts (local, global) and the Euler characteristic
    std::vector<std::string> ShowMe() const
    {
        static const char *name[4] = {"N_elements", "N_facets  ", "N_ridges  ", "N_peaks   "};
        std::vector<std::string> out;
        char buf[96];
        for (int c = 0; c <= nDim_; ++c) { snprintf(buf, sizeof buf, "  %s = %10d%10d", name[c], n_[c], n_[c]); out.push_back(buf); }
        int chi = 0;
        for (int c = nDim_; c >= 0; c -= 2) chi += n_[c];
        for (int c = nDim_ - 1; c >= 0; c -= 2) chi -= n_[c];
        snprintf(buf, sizeof buf, "Euler Characteristic = %10d%10d", chi, chi);
        out.push_back(buf);
        return out;
    }

    /// CoarsenLocalPartitioning(partitioning, check_topology, preserve_material_interfaces, coarsefaces_algo = 0)
    /// (Topology.cpp:685-828): disconnected partitions are split and empty ones removed, then, codimension by
    /// codimension, the agglomerated entities are the minimal intersection sets; with check_topology every stage is
    /// followed by Show / Mark / DeAgglomerate of the bad agglomerated entities (Topology.cpp:728-739, 421-434).
    std::shared_ptr<AgglomeratedTopology> CoarsenLocalPartitioning(const std::vector<int> &partitioning, bool check_topology = false,
                                                                   bool preserve_material_interfaces = false)
    {
        PARELAG_TEST_FOR_EXCEPTION((int)partitioning.size() != n_[0], std::runtime_error,
                                   "CoarsenLocalPartitioning(): partitioning has the wrong size");
        Partition_ = partitioning;
        messages_.clear();
        int nAE = 0;
        if (preserve_material_interfaces)
        {
            // connectedComponents.cpp:90-96: the material-aware form is a stub in the reference
            messages_.push_back("WARNING: this form of connectedComponents not implemented yet.");
            for (int p : Partition_) nAE = std::max(nAE, p + 1);
        }
        else
        {
            Timer t = TimeManager::AddTimer("Mesh Agglomeration: connected components");
            nAE = ConnectedComponents(Partition_, B_[0], hostcsr::Transpose(B_[0]));
        }
        // TransposeOrientation(partitioning, nAE): AE x element, +1
        HostCSR el_AE;
        el_AE.nrows = n_[0]; el_AE.ncols = nAE;
        el_AE.I.resize(n_[0] + 1); std::iota(el_AE.I.begin(), el_AE.I.end(), 0);
        el_AE.J.assign(Partition_.begin(), Partition_.end());
        el_AE.A.assign(n_[0], 1.0);
        AEntity_entity_.clear();
        AEntity_entity_.push_back(hostcsr::Transpose(el_AE));
        if (check_topology) CheckStage(0);
        std::vector<HostCSR> cB;
        for (int icodim = 0; icodim < nDim_; ++icodim)
        {
            HostCSR AE_fc = hostcsr::MultOrientation(AEntity_entity_[icodim], B_[icodim]);
            HostCSR fc_AE = hostcsr::Transpose(AE_fc);
            if (icodim == 0 && HasBdrAttributes())
            {
                // fc_AE_fc + fc_bdrAttr_fc == [AE_fc; bdrAttr^T]^T [AE_fc; bdrAttr^T]: the boundary
                // attributes act as extra agglomerates (columns nAE + attribute)
                HostCSR shifted = facet_bdrAttribute_;
                for (auto &c : shifted.J) c += AE_fc.nrows;
                shifted.ncols += AE_fc.nrows;
                fc_AE.ncols = shifted.ncols;
                fc_AE = hostcsr::Add(fc_AE, shifted);
            }
            HostCSR fc_AF = MinimalIntersectionSetsFromMembership(fc_AE);
            AEntity_entity_.push_back(hostcsr::Transpose(fc_AF));
            if (check_topology && CheckStage(icodim + 1)) fc_AF = hostcsr::Transpose(AEntity_entity_[icodim + 1]);
            cB.push_back(hostcsr::MultOrientation(AE_fc, fc_AF));
        }
        HostCSR cbdr;
        if (HasBdrAttributes()) cbdr = hostcsr::MultOrientation(AEntity_entity_[1], facet_bdrAttribute_);
        auto coarse = std::make_shared<AgglomeratedTopology>(std::move(cB), std::move(cbdr), nDim_);
        CoarserTopology_ = coarse;
        coarse->FinerTopology_ = shared_from_this();
        owned_coarser_ = coarse;
        return coarse;
    }

private:
    /// Topology.cpp:728-739 (elements) and CheckHFacetsTopology (:421-434): report, mark and de-agglomerate the bad
    /// agglomerated entities of one codimension; true when the table changed
    bool CheckStage(int codim)
    {
        using ATC = AgglomeratedTopologyCheck;
        ATC::ShowBadAgglomeratedEntities(nDim_, B_, AEntity_entity_, codim, messages_);
        std::vector<int> isbad;
        if (!ATC::MarkBadAgglomeratedEntities(nDim_, B_, AEntity_entity_, codim, isbad)) return false;
        HostCSR fixed = ATC::DeAgglomerate(AEntity_entity_[codim], isbad);
        messages_.push_back("Correcting agglomerated topology for icodim: " + std::to_string(codim));
        messages_.push_back("  original number of agglomerates: " + std::to_string(AEntity_entity_[codim].nrows));
        messages_.push_back("  number which were bad: " + std::to_string(std::accumulate(isbad.begin(), isbad.end(), 0)));
        messages_.push_back("  number of new agglomerates after de-agglomeration: " + std::to_string(fixed.nrows - AEntity_entity_[codim].nrows));
        AEntity_entity_[codim] = std::move(fixed);
        return true;
    }
    std::vector<std::string> messages_;
    int nDim_;
    std::vector<HostCSR> B_;
    std::vector<int> n_;
    HostCSR facet_bdrAttribute_;
    std::vector<HostCSR> AEntity_entity_;
    std::vector<int> Partition_;
    std::weak_ptr<AgglomeratedTopology> CoarserTopology_, FinerTopology_;
    std::shared_ptr<AgglomeratedTopology> owned_coarser_;
    mutable std::map<std::pair<int, int>, HostCSR> conn_;
};

/// MFEMRefinedMeshPartitioner on the lexicographic structured grid: the agglomerate
/// of a fine element is its parent in one uniform refinement, AE(i,j,k) = (i/2,j/2,k/2)
inline std::vector<int> RefinedHexPartition(int nx, int ny, int nz)
{
    PARELAG_TEST_FOR_EXCEPTION(nx % 2 || ny % 2 || nz % 2, std::runtime_error,
                               "RefinedHexPartition(): the grid is not a uniform refinement of a coarser grid");
    const int cx = nx / 2, cy = ny / 2;
    std::vector<int> p((size_t)nx * ny * nz);
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) p[(size_t)i + (size_t)nx * (j + (size_t)ny * k)] = ((k / 2) * cy + (j / 2)) * cx + (i / 2);
    return p;
}
/// Logical Cartesian agglomeration (LogicalPartitioner::Partition with CoarsenLogicalCartesianOperator,
/// src/partitioning/LogicalPartitioner.hpp:46-103, CartesianPartitioner.hpp:113-131): elements with the same coarse index
/// (i/rx, j/ry, k/rz) form an agglomerate, also when the grid is not a multiple of the ratio (ragged last blocks: the
/// 60 x 220 x 85 SPE10 grid).  The reference numbers the parts in the order its flood fill meets them while scanning the
/// elements in storage order; on the lexicographic grid that is the lexicographic order of the coarse indices.
/// Coarse dimensions: ceil(n / r).
inline std::vector<int> CartesianHexPartition(int nx, int ny, int nz, int rx = 2, int ry = 2, int rz = 2)
{
    PARELAG_TEST_FOR_EXCEPTION(rx < 1 || ry < 1 || rz < 1, std::runtime_error, "CartesianHexPartition(): bad coarsening ratio");
    const int cx = (nx + rx - 1) / rx, cy = (ny + ry - 1) / ry;
    std::vector<int> p((size_t)nx * ny * nz);
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) p[(size_t)i + (size_t)nx * (j + (size_t)ny * k)] = ((k / rz) * cy + (j / ry)) * cx + (i / rx);
    return p;
}
/// GeometricBoxPartitioner::doPartition (src/partitioning/GeometricBoxPartitioner.cpp:20-79): the bounding box is cut into
/// round(extent / r) boxes per direction, r = (volume / num_partitions)^(1/3); an element belongs to the box that holds
/// the mean of its vertices (centroids: n x 3); partition = ix + nx (iy + ny iz), compacted to contiguous ids (boxes that
/// hold no element disappear).
inline std::vector<int> GeometricBoxPartition(const double *centroids, int n, const double *bmin, const double *bmax, int num_partitions)
{
    PARELAG_TEST_FOR_EXCEPTION(num_partitions < 1, std::runtime_error, "GeometricBoxPartition(): bad number of partitions");
    double vol = 1.0, ext[3];
    for (int a = 0; a < 3; ++a) { ext[a] = bmax[a] - bmin[a]; vol *= ext[a]; }
    const double radius = std::pow(vol / (double)num_partitions, 1.0 / 3.0);
    int ndir[3];
    double pr[3];
    for (int a = 0; a < 3; ++a)
    {
        ndir[a] = (int)(ext[a] / radius + 0.5);
        PARELAG_TEST_FOR_EXCEPTION(ndir[a] < 1, std::runtime_error, "GeometricBoxPartition(): degenerate bounding box");
        pr[a] = ext[a] / (double)ndir[a];
    }
    std::vector<int> part((size_t)n);
    for (int e = 0; e < n; ++e)
    {
        int w[3];
        for (int a = 0; a < 3; ++a) w[a] = (int)((centroids[3 * (size_t)e + a] - bmin[a]) / pr[a]);
        part[e] = w[0] + ndir[0] * (w[1] + ndir[1] * w[2]);
    }
    std::vector<int> ids(part);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    for (int e = 0; e < n; ++e) part[e] = (int)(std::lower_bound(ids.begin(), ids.end(), part[e]) - ids.begin());
    return part;
}
/// LogicalPartitioner with LogicalCartesianMaterialId / CoarsenLogicalCartesianOperatorMaterialId
/// (src/partitioning/LogicalPartitioner.hpp:46-128, CartesianPartitioner.hpp:30-36,133-150; examples/LogicalPartitionerDemo.cpp:
/// 203-226): every element carries a logical index (i, j, k) and a material id; a flood fill over the element-element table
/// (elements that share a facet) puts neighbours with the same COARSE logical index (i / rx, j / ry, k / rz, material id)
/// into one partition, partitions numbered in the order a scan of the elements starts them.  Unlike CartesianHexPartition
/// this keeps material ids apart and works level after level on the agglomerated topology.
struct LogicalCartesianMaterialId
{
    int i, j, k, materialId;
    bool operator==(const LogicalCartesianMaterialId &o) const { return i == o.i && j == o.j && k == o.k && materialId == o.materialId; }
};
inline LogicalCartesianMaterialId CoarsenLogical(const LogicalCartesianMaterialId &f, const int *ratio)
{
    return LogicalCartesianMaterialId{f.i / ratio[0], f.j / ratio[1], f.k / ratio[2], f.materialId};
}
inline std::vector<int> LogicalPartition(const HostCSR &el_facet, const std::vector<LogicalCartesianMaterialId> &fine_logical, const int *ratio)
{
    const int n = el_facet.nrows;
    PARELAG_TEST_FOR_EXCEPTION((int)fine_logical.size() != n, std::runtime_error, "LogicalPartition(): one logical index per element");
    const HostCSR facet_el = hostcsr::Transpose(el_facet);
    std::vector<int> part((size_t)n, -1), queue;
    queue.reserve((size_t)n);
    int nparts = 0;
    size_t head = 0;
    for (int e = 0; e < n; ++e)
    {
        if (part[e] >= 0) continue;
        part[e] = nparts++;
        queue.push_back(e);
        for (; head < queue.size(); ++head)
        {
            const int i = queue[head];
            const LogicalCartesianMaterialId ci = CoarsenLogical(fine_logical[i], ratio);
            for (int kf = el_facet.I[i]; kf < el_facet.I[i + 1]; ++kf)
            {
                const int f = el_facet.J[kf];
                for (int ke = facet_el.I[f]; ke < facet_el.I[f + 1]; ++ke)
                {
                    const int k = facet_el.J[ke];
                    if (part[k] < 0 && ci == CoarsenLogical(fine_logical[k], ratio)) { part[k] = part[i]; queue.push_back(k); }
                }
            }
        }
    }
    return part;
}
/// LogicalPartitioner::ComputeCoarseLogical (:105-128): the coarse logical index of an agglomerate is that of its first element
inline std::vector<LogicalCartesianMaterialId> ComputeCoarseLogical(const HostCSR &AE_element, const std::vector<LogicalCartesianMaterialId> &fine_logical,
                                                                    const int *ratio)
{
    std::vector<LogicalCartesianMaterialId> c((size_t)AE_element.nrows);
    for (int a = 0; a < AE_element.nrows; ++a) c[a] = CoarsenLogical(fine_logical[AE_element.J[AE_element.I[a]]], ratio);
    return c;
}
/// Process-wide options of the hierarchy builders (what the reference's drivers take from their command lines:
/// testsuite/UpscalingGeneralForm.cpp --geometric, testsuite/twentyseven.cpp --partition / --check):
/// partitioner 0 = derefinement / logical Cartesian (default), 1 = geometric boxes, 2 = the given element partitioning
/// (1 and 2: two levels), 3 = LogicalPartitioner with material ids (user_partitioning holds the material id of every fine
/// element; examples/LogicalPartitionerDemo.cpp); check_topology = second argument of CoarsenLocalPartitioning; log = the lines the topology
/// coarsening reported since the options were last set.
struct TopologyOptions
{
    int partitioner = 0;
    bool check_topology = false;
    std::vector<int> user_partitioning;
    std::vector<std::string> log;
};
inline TopologyOptions &GlobalTopologyOptions() { static TopologyOptions o; return o; }
/// one coarsening step of the builders: CoarsenLocalPartitioning with the process-wide check flag; the reference's
/// messages go to stdout (SerializedOutput at Topology.cpp:731, 426) and to the log
inline std::shared_ptr<AgglomeratedTopology> CoarsenWithOptions(AgglomeratedTopology &fine, const std::vector<int> &partitioning,
                                                                bool preserve_material_interfaces = false)
{
    TopologyOptions &opt = GlobalTopologyOptions();
    auto coarse = fine.CoarsenLocalPartitioning(partitioning, opt.check_topology, preserve_material_interfaces);
    for (const std::string &line : fine.Messages()) { std::cout << line << std::endl; opt.log.push_back(line); }
    return coarse;
}
} // namespace parelag
