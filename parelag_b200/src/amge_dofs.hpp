// amge_dofs.hpp -- host-side dof bookkeeping of the coarsening path (integer, bit-exact):
//   DofHandlerALG numbering and entity->dof tables   src/amge/DofHandler.cpp:694-1463
//   DofHandlerFE closure tables (lowest order)       src/amge/DofHandler.cpp:347-392
//   DofAgglomeration (AE->dof rows, interior first)  src/amge/DOFAgglomeration.cpp:33-315,503-531
//   ElementalMatricesContainer                       src/amge/ElementalMatricesContainer.cpp:55-212
// plus the per-level data bag the coarsening kernels consume.
#pragma once
#include <map>
#include "amge_topology.hpp"

namespace parelag
{
/// one dense (row-major) matrix per entity: the "DG-like" block-diagonal mass matrices M_[idx]
/// allocator whose construct() default-initialises: resize() of a gigabyte pool does not zero-fill it serially
/// before the (parallel) fill writes every entry anyway
template <class T> struct default_init_allocator : std::allocator<T>
{
    template <class U> struct rebind { using other = default_init_allocator<U>; };
    using std::allocator<T>::allocator;
    template <class U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void *>(p)) U; }
    template <class U, class... Args> void construct(U *p, Args &&...args) { ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...); }
};

struct BlockPool
{
    std::vector<int64_t> off{0};   // value offset of block e (size n+1)
    std::vector<int> size;         // order of block e
    std::vector<double, default_init_allocator<double>> vals;
    int n() const { return (int)size.size(); }
    double *add(int m)
    {
        size.push_back(m);
        vals.resize(vals.size() + (size_t)m * m, 0.0);
        off.push_back((int64_t)vals.size());
        return vals.data() + off[off.size() - 2];
    }
    const double *block(int e) const { return vals.data() + off[e]; }
    /// rDof offset of entity e (= sum of sizes before it)
    std::vector<int> rdof_offsets() const
    {
        std::vector<int> o(size.size() + 1, 0);
        for (size_t e = 0; e < size.size(); ++e) o[e + 1] = o[e] + size[e];
        return o;
    }
};

enum DofType : int8_t { DOF_EMPTY = 0, DOF_RANGET = 1, DOF_NULLSPACE = 2 };

class DofHandlerX : public DofHandler
{
public:
    DofHandlerX(int maxCodimensionBaseForDof, std::shared_ptr<AgglomeratedTopology> topo)
        : mcb(maxCodimensionBaseForDof), topo_(std::move(topo)), entity_dof(mcb + 1), n_rangeT(mcb + 1), n_null(mcb + 1),
          int_offsets(mcb + 1), type_ndofs(mcb + 2, 0)
    {
        for (int c = 0; c <= mcb; ++c)
        {
            n_rangeT[c].assign(topo_->GetNumberLocalEntities(c), 0);
            n_null[c].assign(topo_->GetNumberLocalEntities(c), 0);
        }
    }
    int GetMaxCodimensionBaseForDof() const { return mcb; }
    const HostCSR &GetEntityDofTable(int c) const { return entity_dof.at(c); }
    AgglomeratedTopology &Topology() const { return *topo_; }

    void SetDofType(int dof, DofType t)
    {
        if ((int)dof_type.size() <= dof) dof_type.resize(dof + 1, DOF_EMPTY);
        PARELAG_TEST_FOR_EXCEPTION(dof_type[dof] != DOF_EMPTY, std::runtime_error,
                                   "DofHandlerALG::SetDofType(...): DofType[" << dof << "] is already set");
        dof_type[dof] = t;
    }
    /// computeOffset(c) + build{Peak,Ridge,Facet,Element}DofTable: a row lists the
    /// interior dofs of the entity's lower-dimensional carriers (PEAK first), then its own
    void BuildEntityDofTable(int c)
    {
        const int nent = topo_->GetNumberLocalEntities(c);
        const int start = c < mcb ? type_ndofs[c + 1] : 0;
        PARELAG_TEST_FOR_EXCEPTION(ndofs != start, std::runtime_error, "DofHandlerALG::computeOffset(...): nDofs and entityType_nDofs do not agree");
        auto &offs = int_offsets[c];
        offs.assign(nent + 1, start);
        for (int e = 0; e < nent; ++e)
        {
            const int cnt = n_rangeT[c][e] + n_null[c][e];
            PARELAG_TEST_FOR_EXCEPTION(cnt > 100, std::runtime_error, "DofHandlerALG::computeOffset(...): So many dofs on a coarse entity are impossible!");
            offs[e + 1] = offs[e] + cnt;
        }
        ndofs = offs[nent];
        type_ndofs[c] = ndofs;
        HostCSR &T = entity_dof[c];
        T = HostCSR();
        T.nrows = nent; T.ncols = ndofs;
        T.I.assign(nent + 1, 0);
        for (int e = 0; e < nent; ++e)
        {
            for (int small = mcb; small > c; --small)
            {
                const HostCSR &C = topo_->GetConnectivity(c, small);
                const auto &so = int_offsets[small];
                for (int k = C.I[e]; k < C.I[e + 1]; ++k)
                    for (int d = so[C.J[k]]; d < so[C.J[k] + 1]; ++d) { T.J.push_back(d); T.A.push_back(1.0); }
            }
            for (int d = offs[e]; d < offs[e + 1]; ++d) { T.J.push_back(d); T.A.push_back(1.0); }
            T.I[e + 1] = (int)T.J.size();
        }
        for (int cc = c + 1; cc <= mcb; ++cc) entity_dof[cc].ncols = ndofs;
    }
    void GetInteriorDofs(int c, int e, std::vector<int> &dofs) const
    {
        dofs.clear();
        for (int d = int_offsets[c][e]; d < int_offsets[c][e + 1]; ++d) dofs.push_back(d);
    }
    void GetDofsOnBdr(int c, int e, std::vector<int> &dofs) const
    {
        dofs.clear();
        for (int small = mcb; small > c; --small)
        {
            const HostCSR &C = topo_->GetConnectivity(c, small);
            for (int k = C.I[e]; k < C.I[e + 1]; ++k)
                for (int d = int_offsets[small][C.J[k]]; d < int_offsets[small][C.J[k] + 1]; ++d) dofs.push_back(d);
        }
    }
    void GetTypedInteriorDofs(int c, int e, DofType t, std::vector<int> &dofs) const
    {
        dofs.clear();
        for (int d = int_offsets[c][e]; d < int_offsets[c][e + 1]; ++d) if (dof_type[d] == t) dofs.push_back(d);
    }
    /// bit a of mask[d]: d lies on a facet with boundary attribute a+1 (MarkDofsOnSelectedBndr)
    void ComputeBoundaryMask()
    {
        std::vector<uint32_t> mask(ndofs, 0u);
        if (mcb >= 1 && topo_->HasBdrAttributes())
        {
            const HostCSR &fb = topo_->FacetBdrAttribute();
            const HostCSR &FD = entity_dof[1];
            for (int f = 0; f < fb.nrows; ++f)
                if (fb.I[f + 1] - fb.I[f] == 1)
                    for (int k = FD.I[f]; k < FD.I[f + 1]; ++k) mask[FD.J[k]] |= (1u << fb.J[fb.I[f]]);
        }
        SetBoundaryMask(std::move(mask));
    }

    int mcb;
    std::shared_ptr<AgglomeratedTopology> topo_;
    std::vector<HostCSR> entity_dof;
    int ndofs = 0;
    std::vector<int8_t> dof_type;
    std::vector<std::vector<int>> n_rangeT, n_null, int_offsets;
    std::vector<int> type_ndofs;
};

/// AE -> dof rows (interior dofs first, each part sorted by (separator type, dof id));
/// the ADof index of an agglomerated dof is its position in J[c]
class DofAgglomeration
{
public:
    DofAgglomeration(const std::shared_ptr<AgglomeratedTopology> &topo, const DofHandlerX &dof) : ncod(dof.mcb + 1), I(ncod), J(ncod), nint(ncod), slot(ncod), ent_AE(ncod)
    {
        const int nd = dof.ndofs;
        std::vector<HostCSR> AE_dof(ncod);
        for (int c = 0; c < ncod; ++c)
            AE_dof[c] = hostcsr::Mult(hostcsr::Abs(topo->AEntityEntity(c)), hostcsr::Abs(dof.entity_dof[c]));
        sep.assign(nd, 0);
        for (int c = 1; c < ncod; ++c)
            for (int d : AE_dof[c].J) sep[d] = c;
        for (int c = 0; c < ncod; ++c)
        {
            const HostCSR &A = AE_dof[c];
            I[c] = A.I;
            J[c] = A.J;
            nint[c].assign(A.nrows, 0);
            const bool has_bdr = dof.mcb > c;
            // ADof_rDof: AE-local index of every entity-local dof copy (rDof) of member entities
            const HostCSR &ED = dof.entity_dof[c];
            const HostCSR &AEe = topo->AEntityEntity(c);
            slot[c].assign(ED.J.size(), -1);
            ent_AE[c].assign(ED.nrows, -1);
            // agglomerates are independent (an entity belongs to at most one AE of its codimension):
            // every thread sorts and numbers its own agglomerates with a private scratch map
#pragma omp parallel
            {
                std::vector<int> local(nd, -1);
#pragma omp for schedule(static)
                for (int a = 0; a < A.nrows; ++a)
                {
                    int *b = J[c].data() + A.I[a], *e = J[c].data() + A.I[a + 1];
                    if (has_bdr)
                    {
                        std::sort(b, e, [&](int x, int y) { return sep[x] != sep[y] ? sep[x] < sep[y] : x < y; });
                        int cnt = 0;
                        for (int *p = b; p != e; ++p) cnt += (sep[*p] == c);
                        nint[c][a] = cnt;
                    }
                    else nint[c][a] = (int)(e - b);
                    for (int k = I[c][a]; k < I[c][a + 1]; ++k) local[J[c][k]] = k - I[c][a];
                    for (int k = AEe.I[a]; k < AEe.I[a + 1]; ++k)
                    {
                        const int ent = AEe.J[k];
                        ent_AE[c][ent] = a;
                        for (int r = ED.I[ent]; r < ED.I[ent + 1]; ++r) slot[c][r] = local[ED.J[r]];
                    }
                    for (int k = I[c][a]; k < I[c][a + 1]; ++k) local[J[c][k]] = -1;
                }
            }
        }
    }
    int nAE(int c) const { return (int)I[c].size() - 1; }
    int ncod;
    std::vector<std::vector<int>> I, J, nint;
    std::vector<std::vector<int>> slot;     // per c: rdof -> AE-local dof index (-1: entity in no AE)
    std::vector<std::vector<int>> ent_AE;   // per c: fine entity -> AE (-1 none)
    std::vector<int> sep;
};

/// global key + holders of every local entity of one codimension on one level
struct EntitySharing
{
    std::vector<int64_t> key;
    std::vector<int32_t> sI{0}, sJ;
    void push(int64_t k, const std::vector<int32_t> &ranks)
    {
        key.push_back(k);
        sJ.insert(sJ.end(), ranks.begin(), ranks.end());
        std::sort(sJ.begin() + sI.back(), sJ.end());
        sI.push_back((int32_t)sJ.size());
    }
};

/// everything one level of the sequence owns besides P_/D_ (see parelag_sequence.hpp)
struct SequenceData
{
    std::shared_ptr<AgglomeratedTopology> topo;
    int nforms = 4, jstart = 0;
    std::vector<std::shared_ptr<DofHandlerX>> dof;
    std::map<std::pair<int, int>, BlockPool> M;       // (form, codim) -> entity mass blocks
    std::vector<std::vector<double>> targets;          // per form, column-major ndofs x ntargets
    std::vector<int> ntargets;
    std::vector<double> l2const;
    bool is_fe = false;                                // fine level: geometric PV traces
    std::vector<double> facet_area, ridge_length;
    double svd_tol = 1e-9;
    std::map<std::string, int64_t> stats;
    std::vector<EntitySharing> entity_sharing;   // multi-rank: per codimension (amge_par.hpp)
};
} // namespace parelag
