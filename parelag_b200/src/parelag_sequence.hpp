// parelag_sequence.hpp -- the DeRhamSequence *coarse-operator interface* the solver
// layer consumes (src/amge/DeRhamSequence.hpp:116-468, .cpp:1082-1253): per level and
// form the interpolator P_, the derivative D_, the dof handler's boundary marking, and
// links to the coarser / finer sequence.  Operators are kept as host CSR (they are
// setup-time objects); ComputeTrueP / ComputeTrueD hand the solver layer device ParCSR
// matrices with the essential columns zeroed *in the pattern* exactly as
// SparseMatrix::EliminateCols does in the reference.
//
// One rank <-> one mesh partition <-> one GPU; on a single rank dof == true dof
// (SharingMap is the identity), which is the case implemented in this round.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include "parelag_core.hpp"
#include "par_host.hpp"

namespace parelag
{
struct SequenceData;   // amge_dofs.hpp: topology, dof handlers, mass matrices, targets of one level
struct HostCSR
{
    int nrows = 0, ncols = 0;
    std::vector<int> I, J;
    std::vector<double> A;
    HostCSR() : I(1, 0) {}
    HostCSR(int nr, int nc, const int *i, const int *j, const double *a)
        : nrows(nr), ncols(nc), I(i, i + nr + 1), J(j, j + i[nr]), A(a, a + i[nr]) {}
    int64_t nnz() const { return I.empty() ? 0 : I.back(); }
    pe_parcsr_host View() const
    {
        pe_parcsr_host H{};
        H.global_num_rows = nrows; H.global_num_cols = ncols;
        H.num_rows = nrows; H.num_cols_diag = ncols; H.num_cols_offd = 0;
        H.diag_i = I.data(); H.diag_j = J.data(); H.diag_data = A.data();
        return H;
    }
};

/// the part of DofHandler the solver layer needs (src/amge/DofHandler.cpp:315-333,812-853)
class DofHandler
{
public:
    int GetNDofs() const { return (int)BdrMask_.size(); }
    /// bit (a) of mask[d] set <=> dof d lies on a facet carrying boundary attribute a+1
    void SetBoundaryMask(std::vector<uint32_t> mask) { BdrMask_ = std::move(mask); }
    const std::vector<uint32_t> &GetBoundaryMask() const { return BdrMask_; }
    int MarkDofsOnSelectedBndr(const mfem::Array<int> &bndrAttributesMarker, mfem::Array<int> &dofMarker) const
    {
        PARELAG_TEST_FOR_EXCEPTION(dofMarker.Size() != GetNDofs(), std::runtime_error,
                                   "DofHandler::MarkDofsOnSelectedBndr(...): Incorrect array size!");
        uint32_t sel = 0;
        for (int a = 0; a < bndrAttributesMarker.Size() && a < 32; ++a)
            if (bndrAttributesMarker[a]) sel |= (1u << a);
        int n = 0;
        for (int d = 0; d < GetNDofs(); ++d)
        {
            dofMarker[d] = (BdrMask_[d] & sel) ? 1 : 0;
            n += dofMarker[d];
        }
        return n;
    }
private:
    std::vector<uint32_t> BdrMask_;
};

class DeRhamSequence : public std::enable_shared_from_this<DeRhamSequence>
{
public:
    explicit DeRhamSequence(int nforms) : nForms_(nforms), Dof_(nforms), P_(nforms), D_(nforms), RawDof_(nforms, nullptr), DofTrueDof_(nforms) {}
    virtual ~DeRhamSequence() = default;

    int GetNumberOfForms() const noexcept { return nForms_; }
    /// the reference's short names (DeRhamSequence.hpp:100-160,329-364,458-461)
    int GetNumForms() const noexcept { return nForms_; }
    int GetNumDofs(int jform) const { return GetNumberOfDofs(jform); }
    int GetNumTrueDofs(int jform) const { return GetNumberOfTrueDofs(jform); }
    const HostCSR *GetD(int jform) const { return GetDerivativeOperator(jform); }
    std::shared_ptr<DeRhamSequence> ViewFinerSequence() const { return FinerSequence_.lock(); }
    template <typename... Ts> std::unique_ptr<mfem::HypreParMatrix> ComputeTrueDerivativeOperator(Ts &&...args) const { return ComputeTrueD(std::forward<Ts>(args)...); }
    template <typename... Ts> std::unique_ptr<mfem::HypreParMatrix> ComputeTrueMassOperator(Ts &&...args) const { return ComputeTrueM(std::forward<Ts>(args)...); }
    int GetNumberOfDofs(int jform) const { auto d = GetDofHandler(jform); return d ? d->GetNDofs() : 0; }
    int GetNumberOfTrueDofs(int jform) const
    {
        return DofTrueDof_.at(jform) ? DofTrueDof_[jform]->GetTrueLocalSize() : GetNumberOfDofs(jform);
    }
    /// multi-rank: dof <-> true dof map of a form (DofHandler::GetDofTrueDof) and the host communicator
    void SetDofTrueDof(int jform, std::shared_ptr<par::SharingMap> m) { DofTrueDof_.at(jform) = std::move(m); }
    const par::SharingMap *GetDofTrueDof(int jform) const { return DofTrueDof_.at(jform).get(); }
    void SetComm(const pe_host_comm *c) { Comm_ = c; }
    const pe_host_comm *GetComm() const { return Comm_; }
    bool IsParallel() const { return Comm_ && Comm_->size > 1; }
    /// IgnoreNonLocalRange(range map, A, domain map) (SharingMap.cpp:930-946): the rows this rank owns,
    /// columns in global true numbering, as a device ParCSR matrix
    std::unique_ptr<mfem::HypreParMatrix> IgnoreNonLocalRange(const par::SharingMap &range, const HostCSR &A, const par::SharingMap &domain) const
    {
        pe_parcsr_owned *M = nullptr;
        PE_CALL(pe_par_assemble(Comm_, 1, A.nrows, A.ncols, A.I.data(), A.J.data(), A.A.data(), range.gid.data(), range.owner.data(),
                                domain.gid.data(), domain.owner.data(), range.start, range.start + range.ntrue, range.global,
                                domain.start, domain.start + domain.ntrue, domain.global, &M));
        std::unique_ptr<mfem::HypreParMatrix> out;
        try { out = make_unique<mfem::HypreParMatrix>(*pe_parcsr_owned_view(M)); }
        catch (...) { pe_parcsr_owned_free(M); throw; }
        pe_parcsr_owned_free(M);
        return out;
    }
    /// Assemble(range map, A, domain map) (SharingMap.cpp:975-1011): the contributions of all holders of a shared dof
    /// are summed on its owner (E_r^T A E_d), as a device ParCSR matrix
    std::unique_ptr<mfem::HypreParMatrix> Assemble(const par::SharingMap &range, const HostCSR &A, const par::SharingMap &domain) const
    {
        pe_parcsr_owned *M = nullptr;
        PE_CALL(pe_par_assemble(Comm_, 0, A.nrows, A.ncols, A.I.data(), A.J.data(), A.A.data(), range.gid.data(), range.owner.data(),
                                domain.gid.data(), domain.owner.data(), range.start, range.start + range.ntrue, range.global,
                                domain.start, domain.start + domain.ntrue, domain.global, &M));
        std::unique_ptr<mfem::HypreParMatrix> out;
        try { out = make_unique<mfem::HypreParMatrix>(*pe_parcsr_owned_view(M)); }
        catch (...) { pe_parcsr_owned_free(M); throw; }
        pe_parcsr_owned_free(M);
        return out;
    }
    /// ComputeTrueM(jform) (DeRhamSequence.cpp:1100-1107): Assemble(dofTrueDof, ComputeMassOperator(jform), dofTrueDof)
    std::unique_ptr<mfem::HypreParMatrix> ComputeTrueM(int jform) const
    {
        const HostCSR M = ComputeMassOperator(jform);
        if (IsParallel()) return Assemble(*DofTrueDof_.at(jform), M, *DofTrueDof_.at(jform));
        return make_unique<mfem::HypreParMatrix>(M.View());
    }
    /// Assemble(dofTrueDof(range form), A, dofTrueDof(domain form)) for a local operator between two forms (e.g. B = W D)
    std::unique_ptr<mfem::HypreParMatrix> AssembleTrue(int range_form, const HostCSR &A, int domain_form) const
    {
        if (IsParallel()) return Assemble(*DofTrueDof_.at(range_form), A, *DofTrueDof_.at(domain_form));
        return make_unique<mfem::HypreParMatrix>(A.View());
    }
    DofHandler *GetDofHandler(int jform) const { return RawDof_.at(jform) ? RawDof_[jform] : Dof_.at(jform).get(); }
    void SetDofHandler(int jform, std::unique_ptr<DofHandler> d) { Dof_.at(jform) = std::move(d); }

    /// P_[j] maps the COARSER level's form j to this level (stored on the finer
    /// sequence, DeRhamSequence.hpp:696-704)
    void SetP(int jform, HostCSR P) { P_.at(jform) = std::make_shared<HostCSR>(std::move(P)); }
    void SetD(int jform, HostCSR D) { D_.at(jform) = std::make_shared<HostCSR>(std::move(D)); }
    const HostCSR *GetP(int jform) const { return P_.at(jform).get(); }
    const HostCSR *GetDerivativeOperator(int jform) const { return D_.at(jform).get(); }

    std::shared_ptr<DeRhamSequence> CoarserSequence() const { return CoarserSequence_.lock(); }
    std::shared_ptr<DeRhamSequence> ViewCoarserSequence() const { return CoarserSequence_.lock(); }
    std::shared_ptr<DeRhamSequence> FinerSequence() const { return FinerSequence_.lock(); }
    void SetCoarserSequence(const std::shared_ptr<DeRhamSequence> &c)
    {
        CoarserSequence_ = c;
        if (c) c->FinerSequence_ = shared_from_this();
    }

    /// GetP(jform, ess) + IgnoreNonLocalRange: columns of marked COARSE dofs are zeroed
    /// but kept in the pattern (DeRhamSequence.cpp:1142-1153,1241-1253)
    std::unique_ptr<mfem::HypreParMatrix> ComputeTrueP(int jform, mfem::Array<int> &ess_label) const
    {
        auto coarser = CoarserSequence_.lock();
        PARELAG_ASSERT(coarser);
        PARELAG_TEST_FOR_EXCEPTION(!P_.at(jform), std::runtime_error, "DeRhamSequence::ComputeTrueP(): P_[" << jform << "] is not available");
        HostCSR P = *P_[jform];
        mfem::Array<int> marker(P.ncols);
        marker = 0;
        coarser->GetDofHandler(jform)->MarkDofsOnSelectedBndr(ess_label, marker);
        for (size_t k = 0; k < P.J.size(); ++k)
            if (marker[P.J[k]]) P.A[k] = 0.0;
        if (IsParallel()) return IgnoreNonLocalRange(*DofTrueDof_.at(jform), P, *coarser->DofTrueDof_.at(jform));
        return make_unique<mfem::HypreParMatrix>(P.View());
    }
    std::unique_ptr<mfem::HypreParMatrix> ComputeTrueP(int jform) const
    {
        PARELAG_TEST_FOR_EXCEPTION(!P_.at(jform), std::runtime_error, "DeRhamSequence::ComputeTrueP(): P_[" << jform << "] is not available");
        if (IsParallel()) return IgnoreNonLocalRange(*DofTrueDof_.at(jform), *P_[jform], *CoarserSequence_.lock()->DofTrueDof_.at(jform));
        return make_unique<mfem::HypreParMatrix>(P_[jform]->View());
    }
    /// ComputeDerivativeOperator(jform, ess) + IgnoreNonLocalRange (DeRhamSequence.cpp:1082-1099,1225-1239)
    std::unique_ptr<mfem::HypreParMatrix> ComputeTrueD(int jform, mfem::Array<int> &ess_label) const
    {
        PARELAG_TEST_FOR_EXCEPTION(!D_.at(jform), std::runtime_error, "DeRhamSequence::ComputeTrueD(): D_[" << jform << "] is not available");
        HostCSR D = *D_[jform];
        mfem::Array<int> marker(D.ncols);
        marker = 0;
        GetDofHandler(jform)->MarkDofsOnSelectedBndr(ess_label, marker);
        for (size_t k = 0; k < D.J.size(); ++k)
            if (marker[D.J[k]]) D.A[k] = 0.0;
        if (IsParallel()) return IgnoreNonLocalRange(*DofTrueDof_.at(jform + 1), D, *DofTrueDof_.at(jform));
        return make_unique<mfem::HypreParMatrix>(D.View());
    }
    std::unique_ptr<mfem::HypreParMatrix> ComputeTrueD(int jform) const
    {
        PARELAG_TEST_FOR_EXCEPTION(!D_.at(jform), std::runtime_error, "DeRhamSequence::ComputeTrueD(): D_[" << jform << "] is not available");
        if (IsParallel()) return IgnoreNonLocalRange(*DofTrueDof_.at(jform + 1), *D_[jform], *DofTrueDof_.at(jform));
        return make_unique<mfem::HypreParMatrix>(D_[jform]->View());
    }

    // ---- coarsening (src/amge/DeRhamSequence.cpp:572-692); implemented in amge_coarsen.cpp:
    // integer tables on the host, the per-agglomerate dense work as batched CUDA kernels
    void SetSVDTol(double tol);
    void SetjformStart(int jform);
    std::shared_ptr<DeRhamSequence> Coarsen();
    /// DeRhamSequence::CheckInvariants (DeRhamSequence.cpp:694-970), the identities that involve this level's and the
    /// coarser level's local operators: CheckD (every D_j is there and non-zero, D_{j+1} D_j = 0 to 1e-9), CheckDP
    /// (D_f P_j = P_{j+1} D_c to 1e-6) and CheckCoarseMassMatrix (M_c = P^T M_f P to 1e-6; levels that hold their mass
    /// matrices).  A violation throws std::runtime_error naming the identity, as the reference's assertions do; the return
    /// value is the largest residual met.  Forms below jformStart -- or, on a sequence with supplied operators, forms
    /// whose operators are absent -- are skipped.  (The projector identities of CheckPi are covered by the tests.)
    double CheckInvariants() const;
    double CheckD() const;
    double CheckDP() const;
    double CheckCoarseMassMatrix() const;
    /// ComputeMassOperator(jform): rDof_dof^T M_e rDof_dof (DofHandler.cpp:270-281), host CSR
    HostCSR ComputeMassOperator(int jform) const;
    std::shared_ptr<SequenceData> data;       // null for sequences with externally supplied operators
    const DofHandler *GetDofHandlerRaw(int jform) const { return RawDof_.at(jform); }
    void SetDofHandlerRaw(int jform, DofHandler *d) { RawDof_.at(jform) = d; }

protected:
    int nForms_;
    std::vector<std::unique_ptr<DofHandler>> Dof_;
    std::vector<std::shared_ptr<HostCSR>> P_, D_;
    std::weak_ptr<DeRhamSequence> CoarserSequence_, FinerSequence_;
    std::vector<DofHandler *> RawDof_;              // handlers owned by `data` (Coarsen path)
    std::vector<std::shared_ptr<par::SharingMap>> DofTrueDof_;   // per form; null on a single rank
    const pe_host_comm *Comm_ = nullptr;
};
} // namespace parelag
