// amge_par.hpp -- the multi-rank glue of the coarsening path: which local entities / dofs of a
// box-decomposed structured mesh are shared with which ranks, on every level.
//
// ParElag keeps elements rank-local during coarsening (CoarsenLocalPartitioning,
// src/topology/Topology.cpp:685-828): every rank coarsens its own partition and the levels
// are glued by SharingMaps (src/structures/SharingMap.cpp).  The reference derives the coarse
// entity_trueEntity tables numerically (Pi * e_tE_e * P, SharingMap.cpp:499-524); here the
// identification is combinatorial: a coarse entity is identified across ranks by the smallest
// global key of its fine members, which both holders compute without communication, because
//   * agglomerates never straddle a partition boundary, and
//   * an interface facet carries a pseudo boundary attribute (6 + box face), so the minimal
//     intersection sets split the interface exactly as the neighbour's sets do.
// Owner of a shared item = smallest holding rank (hypre / ParElag convention).
#pragma once
#include "amge_hex.hpp"
#include "par_host.hpp"

namespace parelag
{
/// P[0] x P[1] x P[2] boxes of n[0] x n[1] x n[2] hexahedra; rank = r0 + P0 * (r1 + P1 * r2)
struct BoxDecomposition
{
    int P[3] = {1, 1, 1}, r[3] = {0, 0, 0}, n[3] = {1, 1, 1};
    BoxDecomposition(const int *procs, int rank, int nx, int ny, int nz)
    {
        for (int a = 0; a < 3; ++a) P[a] = procs[a];
        r[0] = rank % P[0]; r[1] = (rank / P[0]) % P[1]; r[2] = rank / (P[0] * P[1]);
        n[0] = nx; n[1] = ny; n[2] = nz;
    }
    int64_t N(int a) const { return (int64_t)n[a] * P[a]; }
    int o(int a) const { return n[a] * r[a]; }
    int nranks() const { return P[0] * P[1] * P[2]; }
    bool interface(int axis, int side) const { return side == 0 ? r[axis] > 0 : r[axis] < P[axis] - 1; }
    /// box coordinates along `axis` that contain the global index g; node = index of a grid plane
    int boxes(int axis, int64_t g, bool node, int out[2]) const
    {
        int cnt = 0;
        const int b = (int)std::min<int64_t>(g / n[axis], P[axis] - 1);
        if (node && g % n[axis] == 0 && g > 0 && g < N(axis)) out[cnt++] = b - 1;
        out[cnt++] = b;
        return cnt;
    }
    /// ranks holding an entity given per axis (global index, is it a node-type index)
    void sharers(const int64_t g[3], const bool node[3], std::vector<int32_t> &out) const
    {
        int bx[2], by[2], bz[2];
        const int cx = boxes(0, g[0], node[0], bx), cy = boxes(1, g[1], node[1], by), cz = boxes(2, g[2], node[2], bz);
        for (int c = 0; c < cz; ++c) for (int b = 0; b < cy; ++b) for (int a = 0; a < cx; ++a)
            out.push_back(bx[a] + P[0] * (by[b] + P[1] * bz[c]));
    }
};

/// fine level: keys are the entity numbers of the undecomposed N0 x N1 x N2 mesh (amge_hex.hpp)
inline std::vector<EntitySharing> FineEntitySharing(const BoxDecomposition &B)
{
    const int nx = B.n[0], ny = B.n[1], nz = B.n[2];
    const int64_t Nx = B.N(0), Ny = B.N(1), Nz = B.N(2);
    const int ox = B.o(0), oy = B.o(1), oz = B.o(2);
    std::vector<EntitySharing> E(4);
    std::vector<int32_t> rk;
    auto add = [&](int codim, int64_t key, int64_t gi, int64_t gj, int64_t gk, bool a, bool b, bool c)
    {
        const int64_t g[3] = {gi, gj, gk};
        const bool node[3] = {a, b, c};
        rk.clear();
        B.sharers(g, node, rk);
        E[codim].push(key, rk);
    };
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)
        add(0, (ox + i) + Nx * ((oy + j) + Ny * (int64_t)(oz + k)), ox + i, oy + j, oz + k, false, false, false);
    const int64_t NFX = (Nx + 1) * Ny * Nz, NFY = Nx * (Ny + 1) * Nz;
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i <= nx; ++i)
        add(1, (ox + i) + (Nx + 1) * ((oy + j) + Ny * (int64_t)(oz + k)), ox + i, oy + j, oz + k, true, false, false);
    for (int k = 0; k < nz; ++k) for (int j = 0; j <= ny; ++j) for (int i = 0; i < nx; ++i)
        add(1, NFX + (ox + i) + Nx * ((oy + j) + (Ny + 1) * (int64_t)(oz + k)), ox + i, oy + j, oz + k, false, true, false);
    for (int k = 0; k <= nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)
        add(1, NFX + NFY + (ox + i) + Nx * ((oy + j) + Ny * (int64_t)(oz + k)), ox + i, oy + j, oz + k, false, false, true);
    const int64_t NEX = Nx * (Ny + 1) * (Nz + 1), NEY = (Nx + 1) * Ny * (Nz + 1);
    for (int k = 0; k <= nz; ++k) for (int j = 0; j <= ny; ++j) for (int i = 0; i < nx; ++i)
        add(2, (ox + i) + Nx * ((oy + j) + (Ny + 1) * (int64_t)(oz + k)), ox + i, oy + j, oz + k, false, true, true);
    for (int k = 0; k <= nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i <= nx; ++i)
        add(2, NEX + (ox + i) + (Nx + 1) * ((oy + j) + Ny * (int64_t)(oz + k)), ox + i, oy + j, oz + k, true, false, true);
    for (int k = 0; k < nz; ++k) for (int j = 0; j <= ny; ++j) for (int i = 0; i <= nx; ++i)
        add(2, NEX + NEY + (ox + i) + (Nx + 1) * ((oy + j) + (Ny + 1) * (int64_t)(oz + k)), ox + i, oy + j, oz + k, true, true, false);
    for (int k = 0; k <= nz; ++k) for (int j = 0; j <= ny; ++j) for (int i = 0; i <= nx; ++i)
        add(3, (ox + i) + (Nx + 1) * ((oy + j) + (Ny + 1) * (int64_t)(oz + k)), ox + i, oy + j, oz + k, true, true, true);
    return E;
}

/// coarse level: key / holders of an agglomerated entity = those of its member with the smallest key
inline std::vector<EntitySharing> CoarseEntitySharing(const AgglomeratedTopology &fine_topo, const std::vector<EntitySharing> &fine)
{
    std::vector<EntitySharing> E(4);
    std::vector<int32_t> rk;
    for (int c = 0; c < 4; ++c)
    {
        const HostCSR &AEe = fine_topo.AEntityEntity(c);
        for (int a = 0; a < AEe.nrows; ++a)
        {
            PARELAG_TEST_FOR_EXCEPTION(AEe.I[a + 1] == AEe.I[a], std::runtime_error, "CoarseEntitySharing: empty agglomerated entity");
            int best = AEe.J[AEe.I[a]];
            for (int k = AEe.I[a]; k < AEe.I[a + 1]; ++k) if (fine[c].key[AEe.J[k]] < fine[c].key[best]) best = AEe.J[k];
            rk.assign(fine[c].sJ.begin() + fine[c].sI[best], fine[c].sJ.begin() + fine[c].sI[best + 1]);
            for (int k = AEe.I[a]; k < AEe.I[a + 1]; ++k)
            {
                const int e = AEe.J[k];
                PARELAG_TEST_FOR_EXCEPTION(fine[c].sI[e + 1] - fine[c].sI[e] != (int)rk.size() ||
                                           !std::equal(rk.begin(), rk.end(), fine[c].sJ.begin() + fine[c].sI[e]), std::runtime_error,
                                           "CoarseEntitySharing: members of an agglomerated entity are held by different rank sets");
            }
            E[c].push(fine[c].key[best], rk);
        }
    }
    return E;
}

/// dof -> true dof map of one form on one level: a dof inherits key and holders from the entity
/// whose interior carries it (codim, entity key, index inside the entity)
inline std::shared_ptr<par::SharingMap> BuildDofSharingMap(const pe_host_comm *comm, const DofHandlerX &dh, const std::vector<EntitySharing> &ent, bool fe_level)
{
    EntitySharing D;
    std::vector<int32_t> rk;
    D.key.reserve(dh.ndofs);
    if (fe_level)
    {
        // lowest-order FE: dof d <-> entity d of codimension mcb
        const EntitySharing &E = ent[dh.mcb];
        PARELAG_ASSERT((int)E.key.size() == dh.ndofs);
        for (int d = 0; d < dh.ndofs; ++d)
        {
            rk.assign(E.sJ.begin() + E.sI[d], E.sJ.begin() + E.sI[d + 1]);
            D.push(((int64_t)dh.mcb << 60) | (E.key[d] << 8), rk);
        }
    }
    else
    {
        std::vector<int64_t> key(dh.ndofs, -1);
        std::vector<std::pair<int, int>> src(dh.ndofs, {-1, -1});
        for (int c = 0; c <= dh.mcb; ++c)
            for (int e = 0; e + 1 < (int)dh.int_offsets[c].size(); ++e)
                for (int d = dh.int_offsets[c][e]; d < dh.int_offsets[c][e + 1]; ++d)
                {
                    PARELAG_TEST_FOR_EXCEPTION(d - dh.int_offsets[c][e] >= 256, std::runtime_error, "BuildDofSharingMap: too many dofs on one entity");
                    key[d] = ((int64_t)c << 60) | (ent[c].key[e] << 8) | (int64_t)(d - dh.int_offsets[c][e]);
                    src[d] = {c, e};
                }
        for (int d = 0; d < dh.ndofs; ++d)
        {
            PARELAG_TEST_FOR_EXCEPTION(src[d].first < 0, std::runtime_error, "BuildDofSharingMap: dof " << d << " belongs to no entity");
            const EntitySharing &E = ent[src[d].first];
            rk.assign(E.sJ.begin() + E.sI[src[d].second], E.sJ.begin() + E.sI[src[d].second + 1]);
            D.push(key[d], rk);
        }
    }
    auto map = std::make_shared<par::SharingMap>();
    map->SetUp(comm, D.key, D.sI, D.sJ);
    return map;
}
} // namespace parelag
