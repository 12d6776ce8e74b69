// amge_topology.hpp -- host-side integer structures of the coarsening path:
//   signed incidence tables (TopologyTable)          src/topology/TopologyTable.cpp:97-139
//   AgglomeratedTopology::CoarsenLocalPartitioning   src/topology/Topology.cpp:685-828
//   findMinimalIntersectionSets                      src/structures/minimalIntersectionSet.cpp:43-130
//   MFEMRefinedMeshPartitioner::Partition            src/partitioning/MFEMRefinedMeshPartitioner.cpp:48-90
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

    /// CoarsenLocalPartitioning(partitioning, check_topology = false,
    /// preserve_material_interfaces = false, coarsefaces_algo = 0)
    std::shared_ptr<AgglomeratedTopology> CoarsenLocalPartitioning(const std::vector<int> &partitioning)
    {
        PARELAG_TEST_FOR_EXCEPTION((int)partitioning.size() != n_[0], std::runtime_error,
                                   "CoarsenLocalPartitioning(): partitioning has the wrong size");
        Partition_ = partitioning;
        int nAE = 0;
        for (int p : partitioning) nAE = std::max(nAE, p + 1);
        // TransposeOrientation(partitioning, nAE): AE x element, +1
        HostCSR el_AE;
        el_AE.nrows = n_[0]; el_AE.ncols = nAE;
        el_AE.I.resize(n_[0] + 1); std::iota(el_AE.I.begin(), el_AE.I.end(), 0);
        el_AE.J.assign(partitioning.begin(), partitioning.end());
        el_AE.A.assign(n_[0], 1.0);
        AEntity_entity_.clear();
        AEntity_entity_.push_back(hostcsr::Transpose(el_AE));
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
} // namespace parelag
