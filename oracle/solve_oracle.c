/*
 * solve_oracle.c -- CPU restatement (TEST INFRASTRUCTURE ONLY) of the arithmetic on
 * ParElag's AMGe solve path.  Nothing under oracle/ is part of the product; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's shared object.
 *
 * The reference (LLNL/parelag) delegates this arithmetic to hypre / MFEM, whose
 * sources are NOT vendored under /root/reference.  Each function
 * below restates the published algorithm of the third-party routine named in its
 * header and cites the ParElag call site that reaches it.
 *
 *   PARITY STATUS: "parity unpinned" for the individual kernels -- the reference
 *   ships no known-answer test for SpMV / relaxation / RAP / PCG.
 *   The end-to-end coarse-space goldens are pinned in oracle/amge.py (tests/test_oracle_goldens.py, tests/test_goldens_cpu.py, tests/test_topology_check_cpu.py).
 *
 * Conventions: int32 indices, FP64 values, CSR (I,J,A); "diag"/"offd" are the two
 * blocks of a hypre ParCSR matrix; x_ext holds the ghost values of x (already
 * exchanged) in offd column order.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* y = alpha*A*x + beta*y, one CSR block, row-wise, columns in stored order.
 * hypre_CSRMatrixMatvec; reached from mfem::HypreParMatrix::Mult at
 * src/linalg/solver_ops/ParELAG_Hierarchy.cpp:193,234. */
void orc_csr_matvec(int n, const int *I, const int *J, const double *A,
                    double alpha, const double *x, double beta, double *y)
{
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = I[i]; k < I[i + 1]; ++k) s += A[k] * x[J[k]];
        y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + alpha * s;
    }
}

/* ParCSR matvec: y = alpha*(diag*x + offd*x_ext) + beta*y.
 * Structure follows the in-tree integer twin
 * src/hypreExtension/hypre_ParCSRMatrixMatvecBoolInt.c:162-199
 * (diag pass, wait for halo, offd pass accumulates with beta=1). */
void orc_parcsr_matvec(int n, const int *dI, const int *dJ, const double *dA,
                       const int *oI, const int *oJ, const double *oA,
                       double alpha, const double *x, const double *x_ext,
                       double beta, double *y)
{
    orc_csr_matvec(n, dI, dJ, dA, alpha, x, beta, y);
    if (oI && oI[n] > 0) orc_csr_matvec(n, oI, oJ, oA, alpha, x_ext, 1.0, y);
}

/* y = alpha*A^T*x + beta*y (one block; A is nrows x ncols).
 * hypre_CSRMatrixMatvecT; reached from P->MultTranspose, Hierarchy.cpp:202 and
 * HiptmairSmoother.cpp:63. */
void orc_csr_matvec_t(int nrows, int ncols, const int *I, const int *J, const double *A,
                      double alpha, const double *x, double beta, double *y)
{
    for (int j = 0; j < ncols; ++j) y[j] = (beta == 0.0 ? 0.0 : beta * y[j]);
    for (int i = 0; i < nrows; ++i) {
        double xi = alpha * x[i];
        for (int k = I[i]; k < I[i + 1]; ++k) y[J[k]] += A[k] * xi;
    }
}

/* hypre_ParCSRComputeL1Norms(A, option, NULL, &l1): option 1 = full l1 row norm
 * (l1-Jacobi), 2 = |a_ii| + sum|offd| (l1-GS, one thread), 4 = truncated variant,
 * 5 = lumped (row sum of the matrix itself; MFEM "Lumped Jacobi" uses A*1).
 * Sign follows the diagonal entry.  Called inside mfem::HypreSmoother::SetOperator,
 * reached from src/linalg/solver_ops/ParELAG_HypreSmootherWrapper.cpp:28-33. */
void orc_l1_norms(int n, const int *dI, const int *dJ, const double *dA,
                  const int *oI, const double *oA, int option, double *l1)
{
    for (int i = 0; i < n; ++i) {
        double d = 0.0, diag = 0.0, off = 0.0, full = 0.0, lump = 0.0;
        for (int k = dI[i]; k < dI[i + 1]; ++k) {
            full += fabs(dA[k]);
            lump += dA[k];
            if (dJ[k] == i) diag = dA[k];
        }
        if (oI) for (int k = oI[i]; k < oI[i + 1]; ++k) { off += fabs(oA[k]); lump += oA[k]; }
        if (option == 1) d = full + off;
        else if (option == 2) d = fabs(diag) + off;
        else if (option == 4) {
            d = fabs(diag) + off;
            if (d <= 4.0 / 3.0 * fabs(diag)) d = fabs(diag);
        } else if (option == 5) d = lump;
        else d = diag; /* option 0 / 6: plain diagonal */
        if (option >= 1 && option <= 4 && diag < 0.0) d = -d;
        l1[i] = d;
    }
}

/* One l1-Jacobi sweep, hypre_ParCSRRelax option 1:
 *   v = w*(f - A u);  u_i += v_i / l1_i.   (x_ext = ghost u, exchanged before) */
void orc_relax_jacobi(int n, const int *dI, const int *dJ, const double *dA,
                      const int *oI, const int *oJ, const double *oA,
                      const double *l1, double weight,
                      const double *f, double *u, const double *u_ext, double *v)
{
    for (int i = 0; i < n; ++i) v[i] = f[i];
    orc_parcsr_matvec(n, dI, dJ, dA, oI, oJ, oA, -weight, u, u_ext, weight, v);
    for (int i = 0; i < n; ++i) u[i] += v[i] / l1[i];
}

/* One symmetric (forward then backward) hybrid Gauss-Seidel sweep,
 * hypre_ParCSRRelax option 2/4 (divide by l1) -- also used for option 6 with
 * l1 = diag(A).  Ghost values u_ext are frozen for both passes (Jacobi across
 * ranks).  `order` is the row visiting order of the forward pass (NULL = natural
 * 0..n-1, i.e. hypre's order); the backward pass visits the reverse.  A multicolour
 * order is a permutation that lists the rows colour by colour.
 * weight==omega==1 is hypre's fast path; the general path follows the c1/c2 form
 * with a frozen copy `uold` of u taken at the start of each pass. */
void orc_relax_gs(int n, const int *dI, const int *dJ, const double *dA,
                  const int *oI, const int *oJ, const double *oA,
                  const double *l1, double weight, double omega,
                  const int *order, const int *rank_of_row,
                  const double *f, double *u, const double *u_ext, double *uold)
{
    int simple = (weight == 1.0 && omega == 1.0);
    double c1 = omega * weight, c2 = omega * (1.0 - weight);
    for (int pass = 0; pass < 2; ++pass) {
        if (!simple) memcpy(uold, u, sizeof(double) * (size_t)n);
        for (int kk = 0; kk < n; ++kk) {
            int k = pass == 0 ? kk : n - 1 - kk;
            int i = order ? order[k] : k;
            if (l1[i] == 0.0) continue;
            double res = f[i];
            if (oI) for (int q = oI[i]; q < oI[i + 1]; ++q) res -= oA[q] * u_ext[oJ[q]];
            if (simple) {
                for (int q = dI[i]; q < dI[i + 1]; ++q) res -= dA[q] * u[dJ[q]];
                u[i] += res / l1[i];
            } else {
                double res0 = 0.0, res2 = 0.0;
                for (int q = dI[i]; q < dI[i + 1]; ++q) {
                    int j = dJ[q];
                    res0 -= dA[q] * u[j];
                    /* rows already visited in this pass */
                    int visited;
                    if (rank_of_row) visited = pass == 0 ? (rank_of_row[j] < rank_of_row[i])
                                                         : (rank_of_row[j] > rank_of_row[i]);
                    else visited = pass == 0 ? (j < i) : (j > i);
                    if (visited) res2 += dA[q] * (uold[j] - u[j]);
                }
                u[i] += (c1 * (res + res0) + c2 * res2) / l1[i];
            }
        }
    }
}

/* Chebyshev coefficients of hypre_ParCSRRelax_Cheby (par_cheby.c), degree
 * `order` in 1..4, interval [lower, upper] with upper = 1.1*max_eig and
 * lower = (upper - min_eig)*fraction + min_eig.  coefs has order entries. */
int orc_cheby_coefs(double max_eig, double min_eig, double fraction, int order, double *coefs)
{
    if (order > 4) order = 4;
    if (order < 1) order = 1;
    int co = order - 1;
    double upper = max_eig * 1.1;
    double lower = (upper - min_eig) * fraction + min_eig;
    double theta = (upper + lower) / 2, delta = (upper - lower) / 2, den;
    switch (co) {
    case 0: coefs[0] = 1.0 / theta; break;
    case 1:
        den = theta * theta + delta * theta;
        coefs[0] = (delta + 2 * theta) / den; coefs[1] = -1.0 / den; break;
    case 2:
        den = 2 * delta * theta * theta - delta * delta * theta - pow(delta, 3) + 2 * pow(theta, 3);
        coefs[0] = (4 * delta * theta - pow(delta, 2) + 6 * pow(theta, 2)) / den;
        coefs[1] = -(2 * delta + 6 * theta) / den; coefs[2] = 2 / den; break;
    case 3:
        den = -(4 * delta * pow(theta, 3) - 3 * pow(delta, 2) * pow(theta, 2)
                - 3 * pow(delta, 3) * theta + 4 * pow(theta, 4));
        coefs[0] = (6 * pow(delta, 2) * theta - 12 * delta * pow(theta, 2)
                    + 3 * pow(delta, 3) - 16 * pow(theta, 3)) / den;
        coefs[1] = (12 * delta * theta - 3 * pow(delta, 2) + 24 * pow(theta, 2)) / den;
        coefs[2] = -(4 * delta + 16 * theta) / den; coefs[3] = 4 / den; break;
    }
    return order;
}

/* One Chebyshev application with D^{-1/2} scaling (hypre_ParCSRRelax_Cheby,
 * scale=1): u += D^{-1/2} p(D^{-1/2} A D^{-1/2}) D^{-1/2} (f - A u), Horner form.
 * Single-rank form (ghosts folded into x_ext callbacks are not needed by tests).
 * work: 5*n doubles. */
void orc_relax_cheby(int n, const int *dI, const int *dJ, const double *dA,
                     const double *coefs, int order,
                     const double *f, double *u, double *work)
{
    double *ds = work, *r = work + n, *uo = work + 2 * n, *v = work + 3 * n;
    int co = order - 1;
    for (int i = 0; i < n; ++i) {
        double d = 0.0;
        for (int k = dI[i]; k < dI[i + 1]; ++k) if (dJ[k] == i) d = dA[k];
        ds[i] = 1.0 / sqrt(d);
    }
    for (int i = 0; i < n; ++i) r[i] = f[i];
    orc_csr_matvec(n, dI, dJ, dA, -1.0, u, 1.0, r);
    for (int i = 0; i < n; ++i) { r[i] *= ds[i]; uo[i] = u[i]; u[i] = r[i] * coefs[co]; }
    double *t = work + 4 * n;
    for (int c = co - 1; c >= 0; --c) {
        for (int i = 0; i < n; ++i) v[i] = ds[i] * u[i];
        orc_csr_matvec(n, dI, dJ, dA, 1.0, v, 0.0, t);
        for (int i = 0; i < n; ++i) u[i] = coefs[c] * r[i] + ds[i] * t[i];
    }
    for (int i = 0; i < n; ++i) u[i] = uo[i] + ds[i] * u[i];
}

/* hypre_ParCSRMatrixFixZeroRows (Hierarchy.cpp:366-371, HiptmairSmootherFactory.cpp:162):
 * rows whose l1 norm (diag+offd) is < eps get a unit diagonal and zeros elsewhere.
 * Returns the number of rows fixed. */
int orc_fix_zero_rows(int n, const int *dI, const int *dJ, double *dA,
                      const int *oI, double *oA)
{
    const double eps = 2.2204460492503131e-16; /* hypre uses machine eps scaled by 1 */
    int nfixed = 0;
    for (int i = 0; i < n; ++i) {
        double l1 = 0.0;
        for (int k = dI[i]; k < dI[i + 1]; ++k) l1 += fabs(dA[k]);
        if (oI) for (int k = oI[i]; k < oI[i + 1]; ++k) l1 += fabs(oA[k]);
        if (l1 < eps) {
            for (int k = dI[i]; k < dI[i + 1]; ++k) dA[k] = (dJ[k] == i) ? 1.0 : 0.0;
            if (oI) for (int k = oI[i]; k < oI[i + 1]; ++k) oA[k] = 0.0;
            ++nfixed;
        }
    }
    return nfixed;
}

/* ---- sparse products (Gustavson two-pass with marker arrays), as
 * hypre_BoomerAMGBuildCoarseOperator / mfem::Mult(SparseMatrix,SparseMatrix):
 * explicit zeros are KEPT; columns are emitted in first-touch order and then
 * sorted ascending here so the pattern can be compared bit-exactly. ---------- */
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

/* symbolic pass: fills CI (n+1) and returns nnz(C). marker: ncolsB ints. */
int orc_spgemm_symbolic(int n, const int *AI, const int *AJ, const int *BI, const int *BJ,
                        int ncolsB, int *CI)
{
    int *marker = (int *)malloc(sizeof(int) * (size_t)(ncolsB > 0 ? ncolsB : 1));
    for (int j = 0; j < ncolsB; ++j) marker[j] = -1;
    int nnz = 0;
    CI[0] = 0;
    for (int i = 0; i < n; ++i) {
        for (int ka = AI[i]; ka < AI[i + 1]; ++ka) {
            int k = AJ[ka];
            for (int kb = BI[k]; kb < BI[k + 1]; ++kb)
                if (marker[BJ[kb]] != i) { marker[BJ[kb]] = i; ++nnz; }
        }
        CI[i + 1] = nnz;
    }
    free(marker);
    return nnz;
}

void orc_spgemm_numeric(int n, const int *AI, const int *AJ, const double *AA,
                        const int *BI, const int *BJ, const double *BA, int ncolsB,
                        const int *CI, int *CJ, double *CA)
{
    int *marker = (int *)malloc(sizeof(int) * (size_t)(ncolsB > 0 ? ncolsB : 1));
    double *acc = (double *)calloc((size_t)(ncolsB > 0 ? ncolsB : 1), sizeof(double));
    for (int j = 0; j < ncolsB; ++j) marker[j] = -1;
    for (int i = 0; i < n; ++i) {
        int pos = CI[i];
        for (int ka = AI[i]; ka < AI[i + 1]; ++ka) {
            int k = AJ[ka];
            double a = AA[ka];
            for (int kb = BI[k]; kb < BI[k + 1]; ++kb) {
                int j = BJ[kb];
                if (marker[j] != i) { marker[j] = i; CJ[pos++] = j; acc[j] = a * BA[kb]; }
                else acc[j] += a * BA[kb];
            }
        }
        qsort(CJ + CI[i], (size_t)(CI[i + 1] - CI[i]), sizeof(int), cmp_int);
        for (int q = CI[i]; q < CI[i + 1]; ++q) CA[q] = acc[CJ[q]];
    }
    free(marker);
    free(acc);
}

/* transpose by counting sort (mfem::Transpose): rows of A^T have ascending columns */
void orc_csr_transpose(int nrows, int ncols, const int *I, const int *J, const double *A,
                       int *TI, int *TJ, double *TA)
{
    for (int j = 0; j <= ncols; ++j) TI[j] = 0;
    for (int k = 0; k < I[nrows]; ++k) TI[J[k] + 1]++;
    for (int j = 0; j < ncols; ++j) TI[j + 1] += TI[j];
    int *next = (int *)malloc(sizeof(int) * (size_t)(ncols > 0 ? ncols : 1));
    for (int j = 0; j < ncols; ++j) next[j] = TI[j];
    for (int i = 0; i < nrows; ++i)
        for (int k = I[i]; k < I[i + 1]; ++k) {
            int p = next[J[k]]++;
            TJ[p] = i;
            TA[p] = A[k];
        }
    free(next);
}

/* hypre's pseudo random numbers (utilities/random.c: Park-Miller minimal standard,
 * a=16807, m=2^31-1, Schrage factorisation) as used by
 * hypre_ParVectorSetRandomValues inside hypre_ParCSRMaxEigEstimateCG. */
void orc_hypre_rand_vector(int n, int seed, double *v)
{
    const int a = 16807, m = 2147483647, q = 127773, r = 2836;
    int s = seed;
    for (int i = 0; i < n; ++i) {
        int low = s % q, high = s / q;
        int test = a * low - r * high;
        s = test > 0 ? test : test + m;
        v[i] = 2.0 * ((double)s / m) - 1.0;
    }
}

/* thread control of the CPU-baseline timing legs (bench.py sets it explicitly: a launcher such as torchrun
 * exports OMP_NUM_THREADS=1) */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}
int orc_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* number of threads a parallel region really gets (cgroup / affinity limits included) */
int orc_probe_threads(void)
{
    int got = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp single
        got = omp_get_num_threads();
    }
#endif
    return got;
}

/* multi-threaded SpMV used only by the CPU-baseline timing leg (rows are
 * independent, result identical to orc_csr_matvec). */
void orc_csr_matvec_mt(int n, const int *I, const int *J, const double *A,
                       double alpha, const double *x, double beta, double *y)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = I[i]; k < I[i + 1]; ++k) s += A[k] * x[J[k]];
        y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + alpha * s;
    }
}

/* Hybrid GS with `nblocks` contiguous row blocks, Gauss-Seidel inside a block and
 * Jacobi (values frozen at sweep start) across blocks: exactly what the reference
 * computes on `nblocks` MPI ranks (off-rank couplings live in offd and use u_ext).
 * Used by the multi-core CPU baseline; l1 must have been computed with the same
 * blocking (orc_l1_norms_blocked). */
void orc_l1_norms_blocked(int n, const int *dI, const int *dJ, const double *dA,
                          int nblocks, double *l1)
{
    for (int b = 0; b < nblocks; ++b) {
        int lo = (int)((long long)n * b / nblocks), hi = (int)((long long)n * (b + 1) / nblocks);
        for (int i = lo; i < hi; ++i) {
            double d = 0.0, diag = 0.0;
            for (int k = dI[i]; k < dI[i + 1]; ++k) {
                int j = dJ[k];
                if (j == i) { diag = dA[k]; d += fabs(dA[k]); }
                else if (j < lo || j >= hi) d += fabs(dA[k]);
            }
            l1[i] = diag < 0 ? -d : d;
        }
    }
}

void orc_relax_gs_blocked(int n, const int *dI, const int *dJ, const double *dA,
                          const double *l1, int nblocks,
                          const double *f, double *u, double *ufrozen)
{
    memcpy(ufrozen, u, sizeof(double) * (size_t)n);
#pragma omp parallel for schedule(static, 1)
    for (int b = 0; b < nblocks; ++b) {
        int lo = (int)((long long)n * b / nblocks), hi = (int)((long long)n * (b + 1) / nblocks);
        for (int pass = 0; pass < 2; ++pass)
            for (int kk = lo; kk < hi; ++kk) {
                int i = pass == 0 ? kk : hi - 1 - (kk - lo);
                if (l1[i] == 0.0) continue;
                double res = f[i];
                for (int q = dI[i]; q < dI[i + 1]; ++q) {
                    int j = dJ[q];
                    res -= dA[q] * ((j >= lo && j < hi) ? u[j] : ufrozen[j]);
                }
                u[i] += res / l1[i];
            }
    }
}

/* Greedy first-fit colouring (rows in natural order, smallest colour not used by an
 * already coloured neighbour of the row pattern): the spec of PE_GS_ORDER_MULTICOLOR.
 * Returns the number of colours. */
int orc_greedy_colors(int n, const int *I, const int *J, int *color)
{
    int ncol = 0, cap = 64;
    int *mark = (int *)malloc(sizeof(int) * (size_t)cap);
    for (int c = 0; c < cap; ++c) mark[c] = -1;
    for (int i = 0; i < n; ++i) color[i] = -1;
    for (int i = 0; i < n; ++i) {
        if (ncol + 1 > cap) {
            mark = (int *)realloc(mark, sizeof(int) * (size_t)(2 * cap));
            for (int c = cap; c < 2 * cap; ++c) mark[c] = -1;
            cap *= 2;
        }
        for (int k = I[i]; k < I[i + 1]; ++k) { int j = J[k]; if (j != i && color[j] >= 0) mark[color[j]] = i; }
        int c = 0;
        while (c < ncol && mark[c] == i) ++c;
        color[i] = c;
        if (c == ncol) ++ncol;
    }
    free(mark);
    return ncol;
}
