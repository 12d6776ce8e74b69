/*
 * parelag_b200_local.h -- C ABI of the batched per-agglomerate kernels (K11-K13) that
 * DeRhamSequence::Coarsen() drives.  All pointers are HOST pointers; the call uploads
 * the batch description, runs one CTA per agglomerate on the GPU and downloads the
 * results.  Replaces, per agglomerate,
 *   SVD_Calculator::ComputeON + Deflate + CochainProjector::CreateDofFunctional
 *     (src/amge/DeRhamSequence.cpp:1818-1926, src/linalg/dense/ParELAG_SVDCalculator.cpp:192-284)
 *   FacetSaddlePoint / RidgePeakSaddlePoint construction + LDLCalculator solves
 *     (src/linalg/solver_core/ParELAG_SaddlePointSolver.cpp:49-189, src/amge/DeRhamSequence.cpp:2364-2556,2779-3027)
 *   CoarsenMassMatrixPart (src/amge/DeRhamSequence.cpp:2139-2166).
 * There is no CPU fallback.
 */
#ifndef PARELAG_B200_LOCAL_H
#define PARELAG_B200_LOCAL_H
#include "parelag_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* one dense row-major matrix per fine entity (ElementalMatricesContainer); rdoff[e] =
 * index of the first "repeated dof" (entity-local dof copy) of entity e */
typedef struct pe_blockpool_view {
    int32_t n;
    const long long *off;      /* n+1 value offsets */
    const int32_t *size;       /* n block orders    */
    const int32_t *rdoff;      /* n+1               */
    const double *vals;
} pe_blockpool_view;

typedef struct pe_csr_view { int32_t nrows, ncols; const int32_t *I, *J; const double *A; } pe_csr_view;

/* a sparse matrix whose rows are written exactly once, in any order (P_ and the coarse
 * D_ while they are being built): row r occupies [start[r], start[r]+len[r]) of J/A */
typedef struct pe_rowpool_view {
    int32_t nrows;
    long long pool_size;
    const long long *start;
    const int32_t *len, *J;
    const double *A;
} pe_rowpool_view;

typedef struct pe_trace_batch {
    int32_t nAE, ndofs;
    const int32_t *I, *J;       /* agglomerated entity -> fine dofs                   */
    const double *pv;           /* PV trace vector (ndofs)                            */
    const double *diagM;        /* agglomerate mass matrix diagonal, per (AE,dof) slot */
    /* agglomerated entities whose mass matrix is NOT diagonal (coarse levels with several dofs per entity): the dense
     * m x m block (row-major, symmetric) starts at denseM[dense_off[ae]]; dense_off[ae] < 0 = diagonal (diagM).  Such an
     * entity takes SVD_Calculator::ComputeON(DenseMatrix&W, ...) (ParELAG_SVDCalculator.cpp:258-284): symmetric
     * eigendecomposition of W (SymEigensolver::ComputeAll), X = W^{1/2}, SVD of X A, back-transform by W^{-1/2}.
     * Both may be NULL when every entity is diagonal. */
    const double *denseM;
    const long long *dense_off;
    int32_t nT, ldT;            /* targets: column-major ldT x nT                     */
    const double *T;
    double svd_tol;             /* SVD_Tolerance_ (relative to pv.M.pv)               */
    const long long *out_off;   /* nAE+1 offsets into out                             */
    /* per AE (m fine dofs, c = nT+1): p[m x c col-major] | mass[c x c] | func[c x m] | sv[nT];
     * only the leading ndofs_out[ae] columns / rows are meaningful (compact leading dimension) */
    double *out;
    int32_t *ndofs_out;         /* 1 + number of retained singular vectors            */
} pe_trace_batch;

typedef struct pe_extension_batch {
    int32_t nAE;
    int32_t facet;              /* 1: hFacetExtension, 0: hRidgePeakExtension          */
    int32_t compute_null;       /* extend the targets and build NullSpace dofs         */
    /* agglomerate -> fine dofs of form j (u), j+1 (p), j+2 (q); interior dofs first   */
    const int32_t *uI, *uJ, *uNint, *pI, *pJ, *pNint, *qI, *qJ;
    const int32_t *aeI, *aeJ;   /* agglomerate -> member fine entities (codim_dom)     */
    pe_blockpool_view Mu, Mp, Mq;        /* entity mass matrices of forms j, j+1, j+2  */
    const int32_t *slot_u, *slot_p, *slot_q;  /* rdof -> agglomerate-local dof index   */
    pe_csr_view Dj, Dj1;        /* fine derivative operators D_j, D_{j+1}              */
    const int32_t *cbI, *cbJ;   /* agglomerate -> coarse dofs on its boundary (ascending) */
    pe_rowpool_view Pj;         /* rows of P_j written by earlier stages               */
    pe_csr_view Pj1;            /* P_{j+1} (complete)                                  */
    pe_rowpool_view Dc;         /* coarse D_j rows written so far (ridge/peak)         */
    const int32_t *pvc;         /* facet: PV coarse dof of form j+1 on the agglomerate */
    const int32_t *pnI, *pnJ;   /* NullSpace coarse dofs of form j+1 on the agglomerate */
    int32_t nT, ldT;
    const double *T;
    double svd_tol, smallest_entry;
    const long long *out_off;
    /* per AE (nu interior u dofs, ncb, nrt = pnI[ae+1]-pnI[ae]):
     * ext[nu x ncb] | bub[nu x nrt] | nul[nu x nT] | lam[ncb] | func[(nrt+nT) x nu] |
     * mass[(ncb+nrt+nT)^2] | sv[nT]   (row-major; k_out[ae] NullSpace dofs retained) */
    double *out;
    int32_t *k_out;
} pe_extension_batch;

int pe_batched_traces(pe_ctx *ctx, const pe_trace_batch *batch);
/* wall-clock split of the pe_batched_extension calls so far (setup profiling): out6 = {H2D staging s, kernel s,
 * D2H s, bytes uploaded, bytes downloaded, calls}; reset != 0 clears the counters */
int pe_local_stage_seconds(double *out6, int reset);
/* enable != 0 opens a scope in which the batched calls keep device copies of their constant inputs (entity mass
 * pools, agglomerate tables, D_j) keyed by host address -- the caller guarantees they do not change or move until
 * the scope is closed with enable == 0, which frees the copies (one scope per DeRhamSequence::Coarsen()). */
int pe_local_cache(pe_ctx *ctx, int enable);
int pe_batched_extension(pe_ctx *ctx, const pe_extension_batch *batch);

#ifdef __cplusplus
}
#endif
#endif
