// amge_coarsen.cpp -- DeRhamSequence::Coarsen() (src/amge/DeRhamSequence.cpp:572-692):
// the host walks the forms (L2 -> H(div) -> H(curl) -> H1) and the stages
//   ComputeCoarseTraces            :1521-2085
//   hFacetExtension                :2214-2581
//   hRidgePeakExtension            :2629-3048
// builds the integer batch description of every stage (agglomerate -> dof lists, boundary
// coarse dofs, ...), hands the per-agglomerate dense work to the batched CUDA kernels of
// csrc/pe_local.cu through the C ABI (include/parelag_b200_local.h), and commits the
// results: coarse dof numbering (DofHandlerALG), P_, the coarse D_, the coarse
// element/facet/ridge mass matrices and the dof functionals of the cochain projector.
// Single rank (dof == true dof).
#include "amge_coarsen.hpp"
#include "amge_par.hpp"
#include "parelag_b200_local.h"
#include <cstring>
#include <malloc.h>
#include <thread>
#include <cstdio>

namespace parelag
{
namespace
{
/// rows written once, in any order (P_ under construction, coarse D_)
struct RowPool
{
    std::vector<long long> start;
    std::vector<int> len, J;
    std::vector<double> A;
    explicit RowPool(int nrows = 0) : start(nrows, 0), len(nrows, 0) {}
    void set_row(int r, const int *cols, const double *vals, int n)
    {
        PARELAG_TEST_FOR_EXCEPTION(len[r] != 0, std::runtime_error, "RowPool: row " << r << " written twice");
        start[r] = (long long)J.size(); len[r] = n;
        J.insert(J.end(), cols, cols + n); A.insert(A.end(), vals, vals + n);
    }
    pe_rowpool_view view() const
    {
        pe_rowpool_view v{};
        v.nrows = (int)start.size(); v.pool_size = (long long)J.size();
        v.start = start.data(); v.len = len.data(); v.J = J.data(); v.A = A.data();
        return v;
    }
    HostCSR to_csr(int ncols) const
    {
        HostCSR C;
        C.nrows = (int)start.size(); C.ncols = ncols;
        C.I.assign(C.nrows + 1, 0);
        for (int r = 0; r < C.nrows; ++r) C.I[r + 1] = C.I[r] + len[r];
        C.J.resize(C.I.back()); C.A.resize(C.I.back());
        for (int r = 0; r < C.nrows; ++r)
        {
            std::copy(J.begin() + start[r], J.begin() + start[r] + len[r], C.J.begin() + C.I[r]);
            std::copy(A.begin() + start[r], A.begin() + start[r] + len[r], C.A.begin() + C.I[r]);
        }
        return C;
    }
};

struct PoolView
{
    std::vector<int> rdoff;
    pe_blockpool_view v{};
    explicit PoolView(const BlockPool &P) : rdoff(P.rdof_offsets())
    {
        v.n = P.n(); v.off = (const long long *)P.off.data(); v.size = P.size.data(); v.rdoff = rdoff.data(); v.vals = P.vals.data();
    }
};
inline pe_csr_view csr_view(const HostCSR &M)
{
    pe_csr_view v{};
    v.nrows = M.nrows; v.ncols = M.ncols; v.I = M.I.data(); v.J = M.J.data(); v.A = M.A.data();
    return v;
}

/// PV trace vector of codimension c (DeRhamSequence3D_FE::computePVTraces on the fine
/// level, DeRhamSequenceAlg::computePVTraces below)
std::vector<double> pv_traces(const SequenceData &S, int c)
{
    const int j = S.nforms - 1 - c;
    const int nd = S.dof[j]->ndofs;
    const HostCSR &AEe = S.topo->AEntityEntity(c);
    std::vector<double> pv(nd, 0.0);
    if (S.is_fe)
    {
        if (c == 0) { std::fill(pv.begin(), pv.end(), 1.0); return pv; }
        for (size_t k = 0; k < AEe.J.size(); ++k)
        {
            const int e = AEe.J[k];
            pv[e] = c == 1 ? AEe.A[k] * S.facet_area[e] : (c == 2 ? AEe.A[k] * S.ridge_length[e] : 1.0);
        }
        return pv;
    }
    const HostCSR &ED = S.dof[j]->entity_dof[c];
    for (size_t k = 0; k < AEe.J.size(); ++k) pv[ED.J[ED.I[AEe.J[k]]]] = AEe.A[k];
    return pv;
}

struct Coarsener
{
    DeRhamSequence &fine;
    SequenceData &S;
    std::shared_ptr<DeRhamSequence> coarse;
    std::shared_ptr<SequenceData> C;
    std::vector<std::unique_ptr<DofAgglomeration>> agg;
    std::vector<RowPool> Ppool;      // per form
    std::vector<RowPool> Dcpool;     // per form j: rows = coarse dofs of form j+1
    std::vector<HostCSR> Pfinal;
    pe_ctx *ctx;

    Coarsener(DeRhamSequence &f) : fine(f), S(*f.data), ctx(Device::Get()) {}

    void run()
    {
        auto ctopo = S.topo->CoarserTopology();
        PARELAG_TEST_FOR_EXCEPTION(!ctopo, std::runtime_error, "DeRhamSequence::Coarsen(): coarsen the topology first");
        const int nf = S.nforms, ndim = nf - 1;
        coarse = std::make_shared<DeRhamSequence>(nf);
        C = std::make_shared<SequenceData>();
        coarse->data = C;
        C->topo = ctopo; C->nforms = nf; C->jstart = S.jstart; C->svd_tol = S.svd_tol; C->is_fe = false;
        C->dof.resize(nf); C->targets.resize(nf); C->ntargets = S.ntargets;
        agg.resize(nf); Ppool.resize(nf); Dcpool.resize(nf); Pfinal.resize(nf);
        // the batched kernels of all forms and codimensions share one device copy of the level's constant inputs
        struct CacheScope
        {
            pe_ctx *c;
            explicit CacheScope(pe_ctx *ctx) : c(ctx) { pe_local_cache(c, 1); }
            ~CacheScope() { pe_local_cache(c, 0); }
        } cache_scope(ctx);
        {
            Timer t = TimeManager::AddTimer("Coarsen: DofAgglomeration");
            for (int j = S.jstart; j < nf; ++j) agg[j] = std::make_unique<DofAgglomeration>(S.topo, *S.dof[j]);
        }
        for (int codim = 0; codim < nf; ++codim)
        {
            const int j = nf - codim - 1;
            if (j < S.jstart) break;
            C->dof[j] = std::make_shared<DofHandlerX>(codim, ctopo);
            Ppool[j] = RowPool(S.dof[j]->ndofs);
            if (j < ndim) Dcpool[j] = RowPool(C->dof[j + 1]->ndofs);
            traces(j);
            if (codim > 0)
            {
                extension(j, nf - j - 2, true);
                if (codim > 1)
                {
                    extension(j, nf - j - 3, false);
                    if (codim > 2) extension(j, nf - j - 4, false);
                }
            }
            {
                Timer t = TimeManager::AddTimer("Coarsen: finalize P and D");
                Pfinal[j] = Ppool[j].to_csr(C->dof[j]->ndofs);
                fine.SetP(j, Pfinal[j]);
                if (codim > 0) coarse->SetD(j, Dcpool[j].to_csr(C->dof[j]->ndofs));
                C->dof[j]->ComputeBoundaryMask();
                coarse->SetDofHandlerRaw(j, C->dof[j].get());
            }
            {
                Timer t = TimeManager::AddTimer("Coarsen: project targets");
                project_targets(j);
            }
        }
        // the constant-one representation in L2
        C->l2const = project(nf - 1, S.l2const, 1);
        fine.SetCoarserSequence(coarse);
    }

    // ------------------------------------------------------------------ traces
    void traces(int j)
    {
        const int codim = S.nforms - 1 - j;
        DofAgglomeration &ag = *agg[j];
        DofHandlerX &cd = *C->dof[j];
        const int nAE = ag.nAE(codim);
        BlockPool &mass = C->M[{j, codim}];
        auto &func = C_func(j, codim);
        func.resize(nAE);
        if (j == 0)
        {
            // Compute0formCoarseTraces: P(vertex, coarse peak) = 1
            for (int a = 0; a < nAE; ++a)
            {
                PARELAG_TEST_FOR_EXCEPTION(ag.I[codim][a + 1] - ag.I[codim][a] != 1, std::runtime_error,
                                           "DeRhamSequence::compute0formCoarseTraces: likely topology error, possibly disconnected edge.");
                const int fd = ag.J[codim][ag.I[codim][a]];
                const double one = 1.0;
                Ppool[0].set_row(fd, &a, &one, 1);
                cd.SetDofType(a, DOF_RANGET);
                cd.n_rangeT[codim][a] = 1;
                func[a] = {1, 1, {1.0}};
                *mass.add(1) = 1.0;
            }
            cd.BuildEntityDofTable(codim);
            return;
        }
        Timer tprep = TimeManager::AddTimer("Coarsen: traces prepare (host)");
        const std::vector<double> pv = pv_traces(S, codim);
        // agglomerate mass matrix M_d: the entities of codimension `codim` carry disjoint dofs, so the block of an
        // agglomerated entity is block diagonal with the blocks of its members.  Diagonal blocks take the row-scaled
        // SVD (SVDCalculator.cpp:247-256); an agglomerated entity with a member block that has a non-zero off-diagonal
        // entry (coarse levels with several dofs per entity: the coarse trace mass is (pv.M.pv) I only up to rounding)
        // takes the dense-weighted SVD (:258-284) like the reference (IsDiagonal, DeRhamSequence.cpp:1857-1868)
        const BlockPool &Me = S.M.at({j, codim});
        const auto rdoff = Me.rdof_offsets();
        std::vector<double> diagM(ag.J[codim].size(), 0.0);
        std::vector<long long> dense_off(nAE, -1);
        std::vector<double> denseM;
        const HostCSR &AEe = S.topo->AEntityEntity(codim);
        for (int a = 0; a < nAE; ++a)
        {
            bool dense = false;
            for (int k = AEe.I[a]; k < AEe.I[a + 1]; ++k)
            {
                const int e = AEe.J[k], m = Me.size[e];
                const double *blk = Me.block(e);
                for (int x = 0; x < m; ++x)
                    for (int y = 0; y < m; ++y)
                    {
                        if (x == y) diagM[ag.I[codim][a] + ag.slot[codim][rdoff[e] + x]] += blk[x * m + x];
                        else if (blk[x * m + y] != 0.0) dense = true;
                    }
            }
            if (!dense) continue;
            const int s0 = ag.I[codim][a], ma = ag.I[codim][a + 1] - s0;
            dense_off[a] = (long long)denseM.size();
            denseM.resize(denseM.size() + (size_t)ma * ma, 0.0);
            double *Md = denseM.data() + dense_off[a];
            for (int k = AEe.I[a]; k < AEe.I[a + 1]; ++k)
            {
                const int e = AEe.J[k], m = Me.size[e];
                const double *blk = Me.block(e);
                for (int x = 0; x < m; ++x)
                    for (int y = 0; y < m; ++y)
                        Md[(size_t)ag.slot[codim][rdoff[e] + x] * ma + ag.slot[codim][rdoff[e] + y]] += blk[x * m + y];
            }
        }
        const int nT = S.ntargets[j];
        std::vector<long long> off(nAE + 1, 0);
        for (int a = 0; a < nAE; ++a)
        {
            const long long m = ag.I[codim][a + 1] - ag.I[codim][a], c = nT + 1;
            off[a + 1] = off[a] + m * c + c * c + c * m + nT;
        }
        std::vector<double> out(off[nAE]);
        std::vector<int> ndofs(nAE);
        pe_trace_batch b{};
        b.nAE = nAE; b.ndofs = S.dof[j]->ndofs; b.I = ag.I[codim].data(); b.J = ag.J[codim].data();
        b.pv = pv.data(); b.diagM = diagM.data(); b.nT = nT;
        if (!denseM.empty()) { b.denseM = denseM.data(); b.dense_off = dense_off.data(); }
        C->stats["trace_dense_mass_" + std::to_string(j)] = (long long)std::count_if(dense_off.begin(), dense_off.end(), [](long long o) { return o >= 0; }); b.ldT = S.dof[j]->ndofs; b.T = S.targets[j].data();
        b.svd_tol = S.svd_tol; b.out_off = off.data(); b.out = out.data(); b.ndofs_out = ndofs.data();
        tprep.Stop();
        {
            Timer t = TimeManager::AddTimer("Coarsen: batched traces (H2D + kernels + D2H)");
            PE_CALL(pe_batched_traces(ctx, &b));
        }
        Timer tcommit = TimeManager::AddTimer("Coarsen: traces commit (host)");
        // commit: dof types / counts, entity table, P rows, coarse trace mass, functionals
        int cnt = 0;
        for (int a = 0; a < nAE; ++a)
        {
            cd.SetDofType(cnt++, DOF_RANGET);
            for (int q = 1; q < ndofs[a]; ++q) cd.SetDofType(cnt++, DOF_NULLSPACE);
            cd.n_rangeT[codim][a] = 1;
            cd.n_null[codim][a] = ndofs[a] - 1;
        }
        cd.BuildEntityDofTable(codim);
        const HostCSR &ED = cd.entity_dof[codim];
        std::vector<double> rowv;
        long long nnull = 0;
        for (int a = 0; a < nAE; ++a)
        {
            const int s = ag.I[codim][a], m = ag.I[codim][a + 1] - s, nc = ndofs[a], cmax = nT + 1;
            const double *p = out.data() + off[a], *ms = p + (size_t)m * cmax, *fn = ms + cmax * cmax;
            nnull += nc - 1;
            rowv.resize(nc);
            for (int i = 0; i < m; ++i)
            {
                for (int c = 0; c < nc; ++c) rowv[c] = p[c * m + i];
                Ppool[j].set_row(ag.J[codim][s + i], ED.J.data() + ED.I[a], rowv.data(), nc);
            }
            double *mb = mass.add(nc);
            std::copy(ms, ms + nc * nc, mb);
            func[a] = {nc, m, std::vector<double>(fn, fn + (size_t)nc * m)};
        }
        C->stats["trace_null_" + std::to_string(j)] = nnull;
    }

    // ------------------------------------------------------------------ extensions
    void extension(int j, int cdom, bool facet)
    {
        Timer tprep = TimeManager::AddTimer("Coarsen: extension prepare (host)");
        const int nf = S.nforms;
        const bool ridge_stuff = (cdom == nf - j - 3);
        DofAgglomeration &au = *agg[j], &ap = *agg[j + 1];
        DofHandlerX &ucd = *C->dof[j], &pcd = *C->dof[j + 1];
        const int nAE = C->topo->GetNumberLocalEntities(cdom);
        const int nT = S.ntargets[j];
        PoolView Mu(S.M.at({j, cdom})), Mp(S.M.at({j + 1, cdom}));
        std::unique_ptr<PoolView> Mq;
        pe_extension_batch b{};
        b.nAE = nAE; b.facet = facet ? 1 : 0; b.compute_null = (facet || ridge_stuff) && nT > 0 ? 1 : 0;
        b.uI = au.I[cdom].data(); b.uJ = au.J[cdom].data(); b.uNint = au.nint[cdom].data();
        b.pI = ap.I[cdom].data(); b.pJ = ap.J[cdom].data(); b.pNint = ap.nint[cdom].data();
        const HostCSR &AEe = S.topo->AEntityEntity(cdom);
        b.aeI = AEe.I.data(); b.aeJ = AEe.J.data();
        b.Mu = Mu.v; b.Mp = Mp.v; b.slot_u = au.slot[cdom].data(); b.slot_p = ap.slot[cdom].data();
        const HostCSR &Dj = *fine.GetDerivativeOperator(j);
        b.Dj = csr_view(Dj);
        // coarse dofs on the boundary of every agglomerate, PV / NullSpace dofs of form j+1
        // two passes (count, fill) over the agglomerates, both thread-parallel; the tables come out in agglomerate
        // order whatever the thread count
        std::vector<int> cbI(nAE + 1, 0), cbJ, pvc(nAE, -1), pnI(nAE + 1, 0), pnJ;
        for (int small = ucd.GetMaxCodimensionBaseForDof(); small > cdom; --small) (void)C->topo->GetConnectivity(cdom, small);   // fill the cache serially
        int bad_pv = -1, unsorted = -1;
#pragma omp parallel
        {
            std::vector<int> tmp;
#pragma omp for schedule(static)
            for (int a = 0; a < nAE; ++a)
            {
                ucd.GetDofsOnBdr(cdom, a, tmp);
                if (!std::is_sorted(tmp.begin(), tmp.end())) unsorted = a;
                cbI[a + 1] = (int)tmp.size();
                if (facet)
                {
                    pcd.GetTypedInteriorDofs(cdom, a, DOF_RANGET, tmp);
                    if (tmp.size() != 1) bad_pv = a; else pvc[a] = tmp[0];
                }
                pcd.GetTypedInteriorDofs(cdom, a, DOF_NULLSPACE, tmp);
                pnI[a + 1] = (int)tmp.size();
            }
        }
        PARELAG_ASSERT(unsorted < 0);
        PARELAG_TEST_FOR_EXCEPTION(bad_pv >= 0, std::runtime_error, "hFacetExtension: expected exactly one PV dof of form " << j + 1 << " per agglomerate");
        for (int a = 0; a < nAE; ++a) { cbI[a + 1] += cbI[a]; pnI[a + 1] += pnI[a]; }
        cbJ.resize((size_t)cbI[nAE]); pnJ.resize((size_t)pnI[nAE]);
#pragma omp parallel
        {
            std::vector<int> tmp;
#pragma omp for schedule(static)
            for (int a = 0; a < nAE; ++a)
            {
                ucd.GetDofsOnBdr(cdom, a, tmp);
                std::copy(tmp.begin(), tmp.end(), cbJ.begin() + cbI[a]);
                pcd.GetTypedInteriorDofs(cdom, a, DOF_NULLSPACE, tmp);
                std::copy(tmp.begin(), tmp.end(), pnJ.begin() + pnI[a]);
            }
        }
        b.cbI = cbI.data(); b.cbJ = cbJ.data(); b.pvc = pvc.data(); b.pnI = pnI.data(); b.pnJ = pnJ.data();
        b.Pj = Ppool[j].view();
        b.Pj1 = csr_view(Pfinal[j + 1]);
        if (!facet)
        {
            DofAgglomeration &aq = *agg[j + 2];
            Mq = std::make_unique<PoolView>(S.M.at({j + 2, cdom}));
            b.qI = aq.I[cdom].data(); b.qJ = aq.J[cdom].data();
            b.Mq = Mq->v; b.slot_q = aq.slot[cdom].data();
            b.Dj1 = csr_view(*fine.GetDerivativeOperator(j + 1));
            b.Dc = Dcpool[j].view();
        }
        b.nT = nT; b.ldT = S.dof[j]->ndofs; b.T = S.targets[j].data();
        b.svd_tol = S.svd_tol; b.smallest_entry = 2.220446049250313e-16;
        std::vector<long long> off(nAE + 1, 0);
        for (int a = 0; a < nAE; ++a)
        {
            const long long nu = au.nint[cdom][a], ncb = cbI[a + 1] - cbI[a], nrt = pnI[a + 1] - pnI[a], lb = ncb + nrt + nT;
            off[a + 1] = off[a] + nu * ncb + nu * nrt + nu * nT + ncb + (nrt + nT) * nu + lb * lb + nT;
        }
        std::vector<double> out(off[nAE]);
        std::vector<int> kout(nAE, 0);
        b.out_off = off.data(); b.out = out.data(); b.k_out = kout.data();
        tprep.Stop();
        {
            Timer t = TimeManager::AddTimer("Coarsen: batched extension (H2D + kernels + D2H)");
            PE_CALL(pe_batched_extension(ctx, &b));
        }
        Timer tcommit = TimeManager::AddTimer("Coarsen: extension commit (host)");
        // ---- commit
        int counter = ucd.ndofs;
        BlockPool &mass = C->M[{j, cdom}];
        auto &func = C_func(j, cdom);
        func.resize(nAE);
        std::vector<int> cols;
        std::vector<double> vals;
        long long tot_rt = 0, tot_null = 0;
        for (int a = 0; a < nAE; ++a)
        {
            const int us = au.I[cdom][a], nu = au.nint[cdom][a];
            const int ncb = cbI[a + 1] - cbI[a], nrt = pnI[a + 1] - pnI[a], k = kout[a];
            const int *cb = cbJ.data() + cbI[a];
            const double *ext = out.data() + off[a], *bub = ext + (size_t)nu * ncb, *nul = bub + (size_t)nu * nrt, *lam = nul + (size_t)nu * nT;
            const double *fn = lam + ncb, *ms = fn + (size_t)(nrt + nT) * nu;
            const int ldm = ncb + nrt + nT, nlb = ncb + nrt + k;
            ucd.n_rangeT[cdom][a] = nrt; ucd.n_null[cdom][a] = k;
            tot_rt += nrt; tot_null += k;
            const int c_rt = counter, c_nu = counter + nrt;
            counter += nrt + k;
            for (int q = 0; q < nrt; ++q) ucd.SetDofType(c_rt + q, DOF_RANGET);
            for (int q = 0; q < k; ++q) ucd.SetDofType(c_nu + q, DOF_NULLSPACE);
            cols.assign(cb, cb + ncb);
            for (int q = 0; q < nrt + k; ++q) cols.push_back(c_rt + q);
            vals.resize(nlb);
            for (int i = 0; i < nu; ++i)
            {
                for (int c = 0; c < ncb; ++c) vals[c] = ext[i * ncb + c];
                for (int q = 0; q < nrt; ++q) vals[ncb + q] = bub[i * nrt + q];
                for (int q = 0; q < k; ++q) vals[ncb + nrt + q] = nul[i * nT + q];
                Ppool[j].set_row(au.J[cdom][us + i], cols.data(), vals.data(), nlb);
            }
            if (facet) Dcpool[j].set_row(pvc[a], cb, lam, ncb);      // coarse exterior derivative of the PV dof
            for (int q = 0; q < nrt; ++q)
            {
                const int col = c_rt + q; const double one = 1.0;
                Dcpool[j].set_row(pnJ[pnI[a] + q], &col, &one, 1);
            }
            double *mb = mass.add(nlb);
            for (int x = 0; x < nlb; ++x) for (int y = 0; y < nlb; ++y) mb[x * nlb + y] = ms[x * ldm + y];
            func[a] = {nrt + k, nu, std::vector<double>(fn, fn + (size_t)(nrt + k) * nu)};
        }
        ucd.BuildEntityDofTable(cdom);
        PARELAG_TEST_FOR_EXCEPTION(ucd.ndofs != counter, std::runtime_error, "extension: coarse dof counter mismatch");
        const std::string tag = (facet ? "facet_ext_" : "ridgepeak_ext_") + std::to_string(j) + "_" + std::to_string(cdom);
        C->stats[tag + "_rangeT"] = tot_rt; C->stats[tag + "_null"] = tot_null;
    }

    // ------------------------------------------------------------------ cochain projector
    struct Rect { int rows = 0, cols = 0; std::vector<double> v; };
    std::map<std::pair<int, int>, std::vector<Rect>> funcs;
    std::vector<Rect> &C_func(int j, int c) { return funcs[{j, c}]; }

    /// CochainProjector::Project (CochainProjector.cpp:147-217) for nv column vectors
    std::vector<double> project(int j, const std::vector<double> &vFine, int nv)
    {
        const int nfd = S.dof[j]->ndofs, ncd = C->dof[j]->ndofs;
        const DofHandlerX &cd = *C->dof[j];
        const DofAgglomeration &ag = *agg[j];
        const HostCSR &P = Pfinal[j];
        std::vector<double> vC((size_t)ncd * nv, 0.0), res(vFine);
        for (int codim = cd.mcb; codim >= 0; --codim)
        {
            const auto &F = funcs[{j, codim}];
            for (int e = 0; e < C->topo->GetNumberLocalEntities(codim); ++e)
            {
                const Rect &f = F[e];
                const int c0 = cd.int_offsets[codim][e];
                const int *fd = ag.J[codim].data() + ag.I[codim][e];
                for (int v = 0; v < nv; ++v)
                    for (int r = 0; r < f.rows; ++r)
                    {
                        double s = 0.0;
                        for (int c = 0; c < f.cols; ++c) s += f.v[(size_t)r * f.cols + c] * res[(size_t)v * nfd + fd[c]];
                        vC[(size_t)v * ncd + c0 + r] = s;
                    }
            }
            res = vFine;
            for (int v = 0; v < nv; ++v)
                for (int i = 0; i < nfd; ++i)
                {
                    double s = 0.0;
                    for (int k = P.I[i]; k < P.I[i + 1]; ++k) s += P.A[k] * vC[(size_t)v * ncd + P.J[k]];
                    res[(size_t)v * nfd + i] -= s;
                }
        }
        return vC;
    }
    void project_targets(int j) { C->targets[j] = project(j, S.targets[j], S.ntargets[j]); }
};
} // namespace

void DeRhamSequence::SetSVDTol(double tol) { PARELAG_ASSERT(data); data->svd_tol = tol; }
void DeRhamSequence::SetjformStart(int jform) { PARELAG_ASSERT(data); data->jstart = jform; }

std::shared_ptr<DeRhamSequence> DeRhamSequence::Coarsen()
{
    PARELAG_TEST_FOR_EXCEPTION(!data, std::runtime_error, "DeRhamSequence::Coarsen(): this sequence holds externally supplied operators only");
    Coarsener c(*this);
    c.run();
    return c.coarse;
}

namespace
{
constexpr double DEFAULT_EQUALITY_TOL = 1e-9, LOOSE_EQUALITY_TOL = 1e-6;       // DeRhamSequence.cpp:36-37
double max_norm(const HostCSR &A) { double m = 0.0; for (double v : A.A) m = std::max(m, std::fabs(v)); return m; }
/// largest |A - B| entry (AreAlmostEqual, ParELAG_MatrixUtils.cpp:33-110); both in canonical CSR form
double max_difference(const HostCSR &A, const HostCSR &B, const std::string &Aname, const std::string &Bname)
{
    PARELAG_TEST_FOR_EXCEPTION(A.nrows != B.nrows || A.ncols != B.ncols, std::logic_error,
                               "AreAlmostEqual(): Size(" << Aname << ") = " << A.nrows << "x" << A.ncols << ", Size(" << Bname << ") = "
                               << B.nrows << "x" << B.ncols << ": sizes don't match!");
    double err = 0.0;
    for (int i = 0; i < A.nrows; ++i)
    {
        int ka = A.I[i], kb = B.I[i];
        while (ka < A.I[i + 1] || kb < B.I[i + 1])
        {
            const int ja = ka < A.I[i + 1] ? A.J[ka] : INT32_MAX, jb = kb < B.I[i + 1] ? B.J[kb] : INT32_MAX;
            if (ja == jb) err = std::max(err, std::fabs(A.A[ka++] - B.A[kb++]));
            else if (ja < jb) err = std::max(err, std::fabs(A.A[ka++]));
            else err = std::max(err, std::fabs(B.A[kb++]));
        }
    }
    return err;
}
} // namespace

double DeRhamSequence::CheckD() const
{
    const int j0 = data ? data->jstart : 0;
    double worst = 0.0;
    for (int j = j0; j < nForms_ - 1; ++j)
    {
        if (!D_[j]) { PARELAG_TEST_FOR_EXCEPTION((bool)data, std::runtime_error, "D_" << j << " is missing"); continue; }
        PARELAG_TEST_FOR_EXCEPTION(D_[j]->nnz() == 0, std::runtime_error, "nnz(D_" << j << ") = 0!");
        PARELAG_TEST_FOR_EXCEPTION(max_norm(*D_[j]) < LOOSE_EQUALITY_TOL, std::runtime_error, "maxNorm(D_" << j << ") = " << max_norm(*D_[j]));
    }
    for (int j = j0; j < nForms_ - 2; ++j)
    {
        if (!D_[j] || !D_[j + 1]) continue;
        const double err = max_norm(hostcsr::Mult(*D_[j + 1], *D_[j]));
        PARELAG_TEST_FOR_EXCEPTION(err > DEFAULT_EQUALITY_TOL, std::runtime_error, "||D_" << j + 1 << " * D_" << j << "|| = " << err);
        worst = std::max(worst, err);
    }
    return worst;
}

double DeRhamSequence::CheckDP() const
{
    auto coarser = CoarserSequence_.lock();
    if (!coarser) return 0.0;
    const int j0 = data ? data->jstart : 0;
    double worst = 0.0;
    for (int j = j0; j < nForms_ - 1; ++j)
    {
        if (!P_[j] || !P_[j + 1] || !D_[j] || !coarser->D_[j]) continue;
        const double err = max_difference(hostcsr::Mult(*D_[j], *P_[j]), hostcsr::Mult(*P_[j + 1], *coarser->D_[j]),
                                          "D_{" + std::to_string(j) + ",fine}*P_" + std::to_string(j),
                                          "P_" + std::to_string(j + 1) + "* D_{" + std::to_string(j) + ",coarse}");
        PARELAG_TEST_FOR_EXCEPTION(err > LOOSE_EQUALITY_TOL, std::runtime_error,
                                   "normInf(D_{" << j << ",fine}*P_" << j << " - P_" << j + 1 << "* D_{" << j << ",coarse}) = " << err);
        worst = std::max(worst, err);
    }
    return worst;
}

double DeRhamSequence::CheckCoarseMassMatrix() const
{
    auto coarser = CoarserSequence_.lock();
    if (!coarser || !data || !coarser->data) return 0.0;
    double worst = 0.0;
    for (int j = data->jstart; j < nForms_; ++j)
    {
        if (!P_[j] || !data->M.count({j, 0}) || !coarser->data->M.count({j, 0})) continue;
        const HostCSR Mrap = hostcsr::Mult(hostcsr::Mult(hostcsr::Transpose(*P_[j]), ComputeMassOperator(j)), *P_[j]);
        const double err = max_difference(coarser->ComputeMassOperator(j), Mrap, "Mcoarse_" + std::to_string(j), "Mrap_" + std::to_string(j));
        PARELAG_TEST_FOR_EXCEPTION(err > LOOSE_EQUALITY_TOL, std::runtime_error, "normInf(Mcoarse_" << j << " - Mrap_" << j << ") = " << err);
        worst = std::max(worst, err);
    }
    return worst;
}

double DeRhamSequence::CheckInvariants() const
{
    // the reference's order (DeRhamSequence.cpp:694-705)
    const double m = CheckCoarseMassMatrix();
    const double d = CheckD();
    const double dp = CheckDP();
    return std::max(m, std::max(d, dp));
}

HostCSR DeRhamSequence::ComputeMassOperator(int jform) const
{
    PARELAG_ASSERT(data);
    const BlockPool &Me = data->M.at({jform, 0});
    const HostCSR &ED = data->dof[jform]->entity_dof[0];
    // assemble sum_e R_e^T M_e R_e with a row-wise accumulator (canonical CSR)
    const int nd = data->dof[jform]->ndofs;
    std::vector<std::vector<std::pair<int, double>>> rows(nd);
    for (int e = 0; e < ED.nrows; ++e)
    {
        const int m = Me.size[e];
        const double *blk = Me.block(e);
        const int *d = ED.J.data() + ED.I[e];
        const double *sg = ED.A.data() + ED.I[e];
        for (int x = 0; x < m; ++x) for (int y = 0; y < m; ++y) rows[d[x]].push_back({d[y], sg[x] * sg[y] * blk[x * m + y]});
    }
    HostCSR M;
    M.nrows = M.ncols = nd; M.I.assign(nd + 1, 0);
    for (int i = 0; i < nd; ++i)
    {
        auto &r = rows[i];
        std::stable_sort(r.begin(), r.end(), [](const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.first < b.first; });
        for (size_t k = 0; k < r.size();)
        {
            double s = 0.0; size_t q = k;
            while (q < r.size() && r[q].first == r[k].first) s += r[q++].second;
            M.J.push_back(r[k].first); M.A.push_back(s);
            k = q;
        }
        M.I[i + 1] = (int)M.J.size();
    }
    return M;
}

std::vector<std::shared_ptr<DeRhamSequence>> BuildHexSequenceHierarchy(int nx, int ny, int nz, double Lx, double Ly, double Lz,
                                                                        const double *alpha, const double *beta, int jstart,
                                                                        int nlevels, double svd_tol, const double *vertex_coords, int beta_components)
{
    return BuildHexSequenceHierarchyPar(nullptr, nullptr, nx, ny, nz, Lx, Ly, Lz, alpha, beta, jstart, nlevels, svd_tol, vertex_coords, beta_components);
}

/// Host memory for the setup phase.  Building a hierarchy allocates (and re-allocates while vectors grow)
/// several KB per fine element; in a container every fresh page costs a serialised page fault, which at
/// 144^3 hexahedra was more than half of the setup time.  ReserveHostArena grows the malloc heap once by
/// `bytes`, touches the pages from all host threads (page faults do run in parallel) and hands the block
/// back to the heap, untrimmed: the setup's std::vectors are then carved out of already-resident memory.
/// mmap-backed malloc is disabled while the arena is in use so that large blocks come from the heap too.
static size_t HostMemAvailable()
{
    size_t avail = (size_t)64 << 30;
    if (FILE *f = fopen("/proc/meminfo", "r"))
    {
        char line[256];
        while (fgets(line, sizeof line, f))
        {
            unsigned long long kb = 0;
            if (sscanf(line, "MemAvailable: %llu kB", &kb) == 1) { avail = (size_t)kb << 10; break; }
        }
        fclose(f);
    }
    return avail;
}

/// `sharers`: processes of this node that reserve an arena at the same time (ranks of the communicator)
void ReserveHostArena(size_t bytes, int sharers)
{
    if (sharers < 1) sharers = 1;
    // never more than a fifth of what the node has left, split between the ranks
    bytes = std::min(bytes, HostMemAvailable() / 5 / (size_t)sharers);
    if (getenv("PE_NO_HOST_ARENA") || bytes < ((size_t)64 << 20)) return;
    // explicit thread count: launchers (torchrun) export OMP_NUM_THREADS=1, which would serialise the first touch
    int nthreads = (int)std::thread::hardware_concurrency() / sharers;
    if (nthreads < 1) nthreads = 1;
    Timer t = TimeManager::AddTimer("Host arena reserve (parallel first touch)");
    if (!getenv("PE_HOST_ARENA_KEEP_MMAP")) mallopt(M_MMAP_MAX, 0);
    mallopt(M_TRIM_THRESHOLD, -1);
    mallopt(M_TOP_PAD, 64 << 20);
    char *p = static_cast<char *>(malloc(bytes));
    if (!p) return;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (size_t k = 0; k < bytes; k += 4096) p[k] = 0;
    free(p);
}
/// back to the default policy (blocks already carved from the heap stay where they are)
void ReleaseHostArena()
{
    if (getenv("PE_NO_HOST_ARENA")) return;
    mallopt(M_MMAP_MAX, 65536);
    mallopt(M_TRIM_THRESHOLD, 128 * 1024);
    mallopt(M_TOP_PAD, 0);
}

std::vector<std::shared_ptr<DeRhamSequence>> BuildHexSequenceHierarchyPar(const pe_host_comm *comm, const int *procs, int nx, int ny, int nz,
                                                                           double Lx, double Ly, double Lz, const double *alpha,
                                                                           const double *beta, int jstart, int nlevels, double svd_tol,
                                                                           const double *vertex_coords, int beta_components)
{
    const bool parallel = comm && comm->size > 1;
    const int one[3] = {1, 1, 1};
    // ~2 KB of entity mass blocks per fine element plus integer tables, interpolation rows and staging buffers
    // (peak RSS of the 144^3 setup: 26 GB = 8.7 KB per element).
    // The heap policy stays in force afterwards (system assembly and BuildSolver stage gigabytes through
    // host vectors as well); ReleaseHostArena() restores glibc's defaults for a caller that wants them back.
    // Sized below the footprint on purpose (4 KB per element, at most 16 GB): measured on the B200 host (64 GB
    // container), a 24 GB arena next to the CUDA driver's own pinned staging pushed the process into reclaim and
    // every cudaMemcpy/cudaMalloc slowed down (setup 40 s), while 12 GB gave 27 s (44 s without an arena).
    // With the arena the heap is never trimmed, so the peak footprint grows from ~9 to ~12 KB per element: ranks
    // that share a node only use it when that still leaves a quarter of the node's memory free.
    const size_t bpe = getenv("PE_HOST_ARENA_BPE") ? (size_t)atoll(getenv("PE_HOST_ARENA_BPE")) : (size_t)4000;
    const size_t nel_all = (size_t)nx * ny * nz, sharers = parallel ? (size_t)comm->size : 1;
    if (nel_all * 12000 * sharers <= HostMemAvailable() / 4 * 3)
        ReserveHostArena(std::min(nel_all * bpe, (size_t)16 << 30), (int)sharers);
    BoxDecomposition box(parallel ? procs : one, parallel ? comm->rank : 0, nx, ny, nz);
    PARELAG_TEST_FOR_EXCEPTION(parallel && box.nranks() != comm->size, std::runtime_error,
                               "BuildHexSequenceHierarchyPar: the process grid does not match the communicator size");
    // Lx, Ly, Lz: extent of THIS rank's box (every box has the same shape)
    StructuredHexMesh mesh(nx, ny, nz, Lx, Ly, Lz);
    if (vertex_coords)
    {
        // multi-rank: every rank passes the vertices of ITS box; copies of an interface vertex must be bitwise
        // equal on the ranks that share it (compute them from the global vertex index), so that the shared facet
        // and ridge mass matrices agree to the last bit, as they do on the reference's ParMesh
        mesh.coords.assign(vertex_coords, vertex_coords + (size_t)3 * mesh.nv());
    }
    if (parallel)
    {
        for (int a = 0; a < 3; ++a) for (int sd = 0; sd < 2; ++sd) mesh.iface[2 * a + sd] = box.interface(a, sd);
        mesh.x0 = Lx * box.r[0]; mesh.y0 = Ly * box.r[1]; mesh.z0 = Lz * box.r[2];
    }
    std::vector<std::shared_ptr<AgglomeratedTopology>> topo(nlevels);
    {
        Timer t = TimeManager::AddTimer("Mesh Agglomeration -- Level 0");
        topo[0] = mesh.Topology();
    }
    int dx = nx, dy = ny, dz = nz;
    const TopologyOptions &topt = GlobalTopologyOptions();
    const bool custom = topt.partitioner != 0;
    if (topt.partitioner == 3)
    {
        // examples/LogicalPartitionerDemo.cpp:203-226: logical Cartesian agglomeration by two per direction that keeps the
        // material ids apart, level after level; CoarsenLocalPartitioning(partitioning, check, preserve_material = 1)
        PARELAG_TEST_FOR_EXCEPTION(parallel, std::runtime_error, "logical partitioner with material ids: single rank");
        const int nel = nx * ny * nz, ratio[3] = {2, 2, 2};
        PARELAG_TEST_FOR_EXCEPTION((int)topt.user_partitioning.size() != nel, std::runtime_error,
                                   "material ids: " << topt.user_partitioning.size() << " entries, the mesh has " << nel << " elements");
        std::vector<LogicalCartesianMaterialId> logical((size_t)nel);
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)
        { const int e = mesh.el(i, j, k); logical[e] = LogicalCartesianMaterialId{i, j, k, topt.user_partitioning[e]}; }
        for (int l = 0; l + 1 < nlevels; ++l)
        {
            Timer t = TimeManager::AddTimer("Mesh Agglomeration -- Level " + std::to_string(l + 1));
            topo[l + 1] = CoarsenWithOptions(*topo[l], LogicalPartition(topo[l]->GetB(0), logical, ratio), true);
            logical = ComputeCoarseLogical(topo[l]->AEntityEntity(0), logical, ratio);
        }
    }
    else if (custom && nlevels > 1)
    {
        PARELAG_TEST_FOR_EXCEPTION(nlevels > 2 || parallel, std::runtime_error, "geometric / user element partitioning: two levels, single rank");
        Timer t = TimeManager::AddTimer("Mesh Agglomeration -- Level 1");
        const int nel = nx * ny * nz;
        if (topt.partitioner == 2)
        {
            // testsuite/twentyseven.cpp:262-288: a hand-made partitioning of the elements
            PARELAG_TEST_FOR_EXCEPTION((int)topt.user_partitioning.size() != nel, std::runtime_error,
                                       "user element partitioning has " << topt.user_partitioning.size() << " entries, the mesh has " << nel << " elements");
            topo[1] = CoarsenWithOptions(*topo[0], topt.user_partitioning);
        }
        else
        {
            // testsuite/UpscalingGeneralForm.cpp:249-256,367-385: one geometric box coarsening, half as many partitions as
            // the once-coarser mesh has elements
            std::vector<double> cen((size_t)3 * nel, 0.0), X;
            if (mesh.deformed()) X = mesh.coords;
            else
            {
                X.resize((size_t)3 * mesh.nv());
                for (int k = 0; k <= nz; ++k) for (int j = 0; j <= ny; ++j) for (int i = 0; i <= nx; ++i)
                { const int v = mesh.vx(i, j, k); X[3 * (size_t)v] = mesh.x0 + i * mesh.hx; X[3 * (size_t)v + 1] = mesh.y0 + j * mesh.hy; X[3 * (size_t)v + 2] = mesh.z0 + k * mesh.hz; }
            }
            double bmin[3] = {1e300, 1e300, 1e300}, bmax[3] = {-1e300, -1e300, -1e300};
            for (size_t v = 0; v < X.size() / 3; ++v) for (int a = 0; a < 3; ++a) { bmin[a] = std::min(bmin[a], X[3 * v + a]); bmax[a] = std::max(bmax[a], X[3 * v + a]); }
            for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)
            {
                double *c = &cen[3 * (size_t)mesh.el(i, j, k)];
                for (int dk = 0; dk < 2; ++dk) for (int dj = 0; dj < 2; ++dj) for (int di = 0; di < 2; ++di)
                    for (int a = 0; a < 3; ++a) c[a] += X[3 * (size_t)mesh.vx(i + di, j + dj, k + dk) + a];
                for (int a = 0; a < 3; ++a) c[a] /= 8.0;
            }
            topo[1] = CoarsenWithOptions(*topo[0], GeometricBoxPartition(cen.data(), nel, bmin, bmax, std::max(1, nel / 8 / 2)));
        }
    }
    for (int l = 0; l + 1 < nlevels && !custom; ++l)
    {
        Timer t = TimeManager::AddTimer("Mesh Agglomeration -- Level " + std::to_string(l + 1));
        // derefinement by two per direction; grids that are not a multiple of two (60 x 220 x 85) take the logical
        // Cartesian agglomeration with ragged last blocks (single rank: the boxes of a decomposition must cut their
        // interfaces in the same way)
        PARELAG_TEST_FOR_EXCEPTION(parallel && (dx % 2 || dy % 2 || dz % 2), std::runtime_error,
                                   "BuildHexSequenceHierarchyPar: box dimensions must be divisible by 2^(levels-1) on more than one rank");
        topo[l + 1] = CoarsenWithOptions(*topo[l], CartesianHexPartition(dx, dy, dz));
        dx = (dx + 1) / 2; dy = (dy + 1) / 2; dz = (dz + 1) / 2;
    }
    std::vector<std::shared_ptr<DeRhamSequence>> seq(nlevels);
    {
        Timer t = TimeManager::AddTimer("DeRhamSequence Construction -- Level 0");
        seq[0] = std::make_shared<DeRhamSequence>(4);
        seq[0]->data = std::make_shared<SequenceData>();
        std::vector<HostCSR> D;
        BuildFineHexSequence(mesh, topo[0], alpha, beta, jstart, *seq[0]->data, D, beta_components);
        for (int j = 0; j < 3; ++j) seq[0]->SetD(j, D[j]);
        for (int j = 0; j < 4; ++j) seq[0]->SetDofHandlerRaw(j, seq[0]->data->dof[j].get());
    }
    if (svd_tol < 0.0)
    {
        // topology-only mode (host integer tables; used by the CPU tests): coarser levels carry
        // their topology but no sequence
        for (int l = 1; l < nlevels; ++l)
        {
            seq[l] = std::make_shared<DeRhamSequence>(4);
            seq[l]->data = std::make_shared<SequenceData>();
            seq[l]->data->topo = topo[l];
        }
    }
    else
        for (int l = 0; l + 1 < nlevels; ++l)
        {
            Timer t = TimeManager::AddTimer("DeRhamSequence Construction -- Level " + std::to_string(l + 1));
            seq[l]->SetSVDTol(svd_tol);
            seq[l + 1] = seq[l]->Coarsen();
        }
    if (parallel)
    {
        // entity and dof sharing on every level (SharingMap); levels without dof handlers
        // (topology-only mode) get the entity tables only
        Timer t = TimeManager::AddTimer("SharingMap construction");
        std::vector<EntitySharing> ent = FineEntitySharing(box);
        for (int l = 0; l < nlevels; ++l)
        {
            if (l > 0) ent = CoarseEntitySharing(*topo[l - 1], ent);
            seq[l]->SetComm(comm);
            seq[l]->data->entity_sharing = ent;
            if (svd_tol < 0.0 && l > 0) continue;
            for (int j = jstart; j < 4; ++j)
                seq[l]->SetDofTrueDof(j, BuildDofSharingMap(comm, *seq[l]->data->dof[j], ent, l == 0));
        }
    }
    return seq;
}
std::vector<std::shared_ptr<DeRhamSequence>> BuildTetSequenceHierarchy(const TetMesh &coarse_mesh, int nref, int nlevels, const double *alpha,
                                                                        const double *beta, int jstart, double svd_tol)
{
    // derefinement agglomeration needs one refinement per coarsening; the geometric / given partitioners do not
    PARELAG_TEST_FOR_EXCEPTION(nlevels < 1 || (GlobalTopologyOptions().partitioner == 0 && nlevels - 1 > nref), std::runtime_error,
                               "BuildTetSequenceHierarchy: " << nlevels << " levels need at least " << nlevels - 1 << " refinements");
    TetMesh mesh = coarse_mesh;
    {
        Timer t = TimeManager::AddTimer("Mesh refinement");
        for (int r = 0; r < nref; ++r) mesh = mesh.Refine();
    }
    const size_t nel_all = (size_t)mesh.nel();
    if (nel_all * 6000 <= HostMemAvailable() / 4 * 3) ReserveHostArena(std::min(nel_all * (size_t)2000, (size_t)16 << 30), 1);
    std::vector<std::shared_ptr<AgglomeratedTopology>> topo(nlevels);
    {
        Timer t = TimeManager::AddTimer("Mesh Agglomeration -- Level 0");
        topo[0] = mesh.Topology();
    }
    int n = mesh.nel();
    const TopologyOptions &topt = GlobalTopologyOptions();
    const bool custom = topt.partitioner != 0;
    if (custom && nlevels > 1)
    {
        PARELAG_TEST_FOR_EXCEPTION(nlevels > 2, std::runtime_error, "geometric / user element partitioning: two levels");
        Timer t = TimeManager::AddTimer("Mesh Agglomeration -- Level 1");
        if (topt.partitioner == 2)
        {
            PARELAG_TEST_FOR_EXCEPTION((int)topt.user_partitioning.size() != n, std::runtime_error,
                                       "user element partitioning has " << topt.user_partitioning.size() << " entries, the mesh has " << n << " elements");
            topo[1] = CoarsenWithOptions(*topo[0], topt.user_partitioning);
        }
        else
        {
            std::vector<double> cen((size_t)3 * n, 0.0);
            double bmin[3] = {1e300, 1e300, 1e300}, bmax[3] = {-1e300, -1e300, -1e300};
            for (int v = 0; v < mesh.nv(); ++v) for (int a = 0; a < 3; ++a) { bmin[a] = std::min(bmin[a], mesh.V[3 * (size_t)v + a]); bmax[a] = std::max(bmax[a], mesh.V[3 * (size_t)v + a]); }
            for (int e = 0; e < n; ++e)
            {
                for (int q = 0; q < 4; ++q) for (int a = 0; a < 3; ++a) cen[3 * (size_t)e + a] += mesh.V[3 * (size_t)mesh.T[4 * e + q] + a];
                for (int a = 0; a < 3; ++a) cen[3 * (size_t)e + a] /= 4.0;
            }
            topo[1] = CoarsenWithOptions(*topo[0], GeometricBoxPartition(cen.data(), n, bmin, bmax, std::max(1, n / 8 / 2)));
        }
    }
    for (int l = 0; l + 1 < nlevels && !custom; ++l)
    {
        Timer t = TimeManager::AddTimer("Mesh Agglomeration -- Level " + std::to_string(l + 1));
        std::vector<int> part((size_t)n);
        for (int e = 0; e < n; ++e) part[e] = e / 8;
        topo[l + 1] = CoarsenWithOptions(*topo[l], part);
        n /= 8;
    }
    std::vector<std::shared_ptr<DeRhamSequence>> seq(nlevels);
    {
        Timer t = TimeManager::AddTimer("DeRhamSequence Construction -- Level 0");
        seq[0] = std::make_shared<DeRhamSequence>(4);
        seq[0]->data = std::make_shared<SequenceData>();
        std::vector<HostCSR> D;
        BuildFineTetSequence(mesh, topo[0], alpha, beta, jstart, *seq[0]->data, D);
        for (int j = 0; j < 3; ++j) seq[0]->SetD(j, D[j]);
        for (int j = 0; j < 4; ++j) seq[0]->SetDofHandlerRaw(j, seq[0]->data->dof[j].get());
    }
    if (svd_tol < 0.0)
        for (int l = 1; l < nlevels; ++l)
        {
            seq[l] = std::make_shared<DeRhamSequence>(4);
            seq[l]->data = std::make_shared<SequenceData>();
            seq[l]->data->topo = topo[l];
        }
    else
        for (int l = 0; l + 1 < nlevels; ++l)
        {
            Timer t = TimeManager::AddTimer("DeRhamSequence Construction -- Level " + std::to_string(l + 1));
            seq[l]->SetSVDTol(svd_tol);
            seq[l + 1] = seq[l]->Coarsen();
        }
    return seq;
}
} // namespace parelag
