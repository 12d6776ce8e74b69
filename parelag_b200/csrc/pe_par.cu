// pe_par.cu -- device side of the multi-rank path: ParCSR transpose-SpMV with the reverse halo
// exchange (hypre_ParCSRMatrixMatvecT), the distributed Galerkin product R^T A P
// (hypre_BoomerAMGBuildCoarseOperator behind mfem::RAP, Hierarchy.cpp:365,400-544), the distributed product A B
// (hypre_ParMatmul behind mfem::ParMult), transpose and a A + b B (hypre_ParCSRMatrixAdd2).
//
// Distributed RAP = local device SpGEMMs on an extended index space + one row exchange:
//   1. fetch the rows of P that belong to A's ghost columns (neighbour exchange, host comm);
//   2. P^ = [P_diag | P_offd] stacked on the fetched rows, columns = own coarse dofs followed by
//      the union G of all ghost coarse dofs;  A^ = [A_diag | A_offd];
//   3. C^ = (P^ restricted to own fine rows)^T (A^ P^)  -- two hash SpGEMMs on the device;
//   4. rows of C^ that belong to ghost coarse dofs are contributions to other ranks' rows:
//      Assemble (pe_par_assemble) sends them to their owners, sums, splits diag / offd and builds
//      the comm package.
#include "pe_core.cuh"
#include "../../include/parelag_b200_par.h"
#include <algorithm>
#include <cub/cub.cuh>
#include <cstring>

extern "C" int pe_ctx_set_host_comm(pe_ctx *ctx, const pe_host_comm *comm)
{
    PE_CHECK(ctx, "null context");
    ctx->hcomm = comm;
    return pe_p2p_init(ctx);      // collective: maps the ranks' halo arenas into each other (NVLink peer memory)
}

// ---------------------------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------------------------
__global__ void k_sel_rowlen(int nsel, const int *__restrict__ rows, const int *__restrict__ dI, const int *__restrict__ oI, int *len)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nsel) return;
    const int r = rows ? rows[k] : k;
    int l = dI[r + 1] - dI[r];
    if (oI) l += oI[r + 1] - oI[r];
    len[k + 1] = l;
    if (k == 0) len[0] = 0;
}
// merged rows: diag entries keep their column, offd entry j becomes omap[j] (or obase + j)
__global__ void k_sel_fill(int nsel, const int *__restrict__ rows, const int *__restrict__ dI, const int *__restrict__ dJ,
                           const double *__restrict__ dA, const int *__restrict__ oI, const int *__restrict__ oJ,
                           const double *__restrict__ oA, const int *__restrict__ omap, int obase,
                           const int *__restrict__ I, int *J, double *A)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nsel) return;
    const int r = rows ? rows[k] : k;
    int q = I[k];
    for (int t = dI[r]; t < dI[r + 1]; ++t, ++q) { J[q] = dJ[t]; A[q] = dA[t]; }
    if (oI) for (int t = oI[r]; t < oI[r + 1]; ++t, ++q) { J[q] = omap ? omap[oJ[t]] : obase + oJ[t]; A[q] = oA[t]; }
}

// out = rows `rows_d[0..nsel)` (all rows when rows_d == null, nsel = nrows) of [diag | offd], offd columns remapped
static int merge_rows(pe_ctx *ctx, const DevCSR &diag, const DevCSR &offd, const int *rows_d, int nsel, const int *omap_d,
                      int obase, int ncols_out, DevCSR &out)
{
    cudaStream_t st = ctx->stream;
    const int *oI = offd.nnz > 0 ? offd.I : nullptr;
    out = DevCSR();
    out.nrows = nsel; out.ncols = ncols_out;
    PE_CUDA(cudaMalloc(&out.I, sizeof(int) * (size_t)(nsel + 1)));
    PE_CUDA(cudaMemsetAsync(out.I, 0, sizeof(int) * (size_t)(nsel + 1), st));
    int nnz = 0;
    if (nsel > 0)
    {
        k_sel_rowlen<<<pe_grid_for(nsel, 256), 256, 0, st>>>(nsel, rows_d, diag.I, oI, out.I); PE_LAUNCHED(ctx);
        void *tmp = nullptr; size_t tb = 0;
        cub::DeviceScan::InclusiveSum(nullptr, tb, out.I, out.I, nsel + 1, st);
        PE_CUDA(cudaMalloc(&tmp, tb));
        PE_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb, out.I, out.I, nsel + 1, st));
        ctx->launches++;
        PE_CUDA(cudaMemcpyAsync(&nnz, out.I + nsel, sizeof(int), cudaMemcpyDeviceToHost, st));
        PE_CUDA(cudaStreamSynchronize(st));
        cudaFree(tmp);
    }
    out.nnz = nnz;
    PE_CUDA(cudaMalloc(&out.J, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
    PE_CUDA(cudaMalloc(&out.A, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
    if (nsel > 0 && nnz > 0)
    {
        k_sel_fill<<<pe_grid_for(nsel, 256), 256, 0, st>>>(nsel, rows_d, diag.I, diag.J, diag.A, oI, offd.J, offd.A, omap_d, obase,
                                                          out.I, out.J, out.A);
        PE_LAUNCHED(ctx);
    }
    return 0;
}

static int download_csr(pe_ctx *ctx, const DevCSR &M, std::vector<int32_t> &I, std::vector<int32_t> &J, std::vector<double> &A)
{
    I.resize((size_t)M.nrows + 1); J.resize((size_t)M.nnz); A.resize((size_t)M.nnz);
    PE_CUDA(cudaMemcpyAsync(I.data(), M.I, sizeof(int32_t) * I.size(), cudaMemcpyDeviceToHost, ctx->stream));
    if (M.nnz)
    {
        PE_CUDA(cudaMemcpyAsync(J.data(), M.J, sizeof(int32_t) * J.size(), cudaMemcpyDeviceToHost, ctx->stream));
        PE_CUDA(cudaMemcpyAsync(A.data(), M.A, sizeof(double) * A.size(), cudaMemcpyDeviceToHost, ctx->stream));
    }
    PE_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// host exchange helper (same protocol as par_host.hpp; duplicated here to keep csrc/ self-contained)
static int exchange(const pe_host_comm *hc, const std::vector<std::vector<char>> &send, std::vector<std::vector<char>> &recv)
{
    const int np = hc->size, me = hc->rank;
    std::vector<int64_t> cnt(np), all((size_t)np * np), sb(np), sd(np), rb(np), rd(np);
    for (int r = 0; r < np; ++r) cnt[r] = (int64_t)send[r].size();
    PE_CHECK(hc->allgather(hc->user, cnt.data(), (int64_t)(sizeof(int64_t) * np), all.data()) == 0, "host communicator: allgather failed");
    int64_t stot = 0, rtot = 0;
    for (int r = 0; r < np; ++r) { sb[r] = cnt[r]; sd[r] = stot; stot += sb[r]; rb[r] = all[(size_t)r * np + me]; rd[r] = rtot; rtot += rb[r]; }
    std::vector<char> sbuf((size_t)std::max<int64_t>(stot, 1)), rbuf((size_t)std::max<int64_t>(rtot, 1));
    for (int r = 0; r < np; ++r) if (sb[r]) memcpy(sbuf.data() + sd[r], send[r].data(), (size_t)sb[r]);
    PE_CHECK(hc->alltoallv(hc->user, sbuf.data(), sb.data(), sd.data(), rbuf.data(), rb.data(), rd.data()) == 0, "host communicator: alltoallv failed");
    recv.assign(np, {});
    for (int r = 0; r < np; ++r) recv[r].assign(rbuf.begin() + rd[r], rbuf.begin() + rd[r] + rb[r]);
    return 0;
}
template <class T> static void put(std::vector<char> &b, const T &v) { const char *c = (const char *)&v; b.insert(b.end(), c, c + sizeof(T)); }

// ---------------------------------------------------------------------------------------------
// extended product  A^ B^  (the common first half of hypre_ParMatmul and of the Galerkin product):
//   rows of B that belong to A's ghost columns are fetched from their owners; B^ = [own rows ; fetched rows] with the
//   columns [own columns of B | G], G = sorted union of all ghost columns; A^ = [A_diag | A_offd]
// ---------------------------------------------------------------------------------------------
struct ExtProduct
{
    DevCSR AB;                       // nA x (ncl + |G|)
    std::vector<int64_t> G, cstart;  // ghost columns (global ids, ascending); column ownership ranges of B
    int ncl = 0;
    int64_t c0 = 0, c1 = 0;
    int nce() const { return ncl + (int)G.size(); }
    void col_ids(int me, std::vector<int64_t> &gid, std::vector<int32_t> &own) const
    {
        gid.resize((size_t)nce()); own.assign((size_t)nce(), me);
        for (int j = 0; j < ncl; ++j) gid[j] = c0 + j;
        for (size_t j = 0; j < G.size(); ++j)
        {
            gid[ncl + j] = G[j];
            own[ncl + j] = (int32_t)(std::upper_bound(cstart.begin(), cstart.end(), G[j]) - cstart.begin()) - 1;
        }
    }
};

static int col_partition(const pe_host_comm *hc, const pe_mat *M, std::vector<int64_t> &start)
{
    start.assign((size_t)hc->size + 1, 0);
    int64_t mine = M->first_col_diag;
    PE_CHECK(hc->allgather(hc->user, &mine, (int64_t)sizeof(int64_t), start.data()) == 0, "host communicator: allgather failed");
    start[hc->size] = M->global_num_cols;
    return 0;
}

static int ext_product(pe_ctx *ctx, const pe_mat *A, const pe_mat *B, ExtProduct &X)
{
    const pe_host_comm *hc = ctx->hcomm;
    PE_CHECK(hc, "distributed sparse products need a host communicator (pe_ctx_set_host_comm)");
    PE_CHECK(A->diag.ncols == B->diag.nrows, "distributed product: size mismatch");
    cudaStream_t st = ctx->stream;
    const int np = hc->size;
    const int nA = A->diag.nrows, nB = B->diag.nrows, ngA = A->offd.ncols, ncl = B->diag.ncols;
    X.ncl = ncl; X.c0 = B->first_col_diag; X.c1 = X.c0 + ncl;
    PE_TRY(col_partition(hc, B, X.cstart));
    // 1. rows of B the neighbours need (their ghost columns of A = my send_map_elmts), as (len, [gcol, val]...)
    std::vector<std::vector<char>> send(np), recv;
    {
        const int nsend = A->send_map_starts.empty() ? 0 : A->send_map_starts.back();
        DevCSR sel;
        std::vector<int32_t> I, J; std::vector<double> V;
        if (nsend > 0)
        {
            PE_TRY(merge_rows(ctx, B->diag, B->offd, A->send_map_d, nsend, nullptr, ncl, ncl + B->offd.ncols, sel));
            PE_TRY(download_csr(ctx, sel, I, J, V));
            devcsr_free(sel);
        }
        for (size_t s = 0; s < A->send_procs.size(); ++s)
        {
            std::vector<char> &b = send[A->send_procs[s]];
            for (int k = A->send_map_starts[s]; k < A->send_map_starts[s + 1]; ++k)
            {
                put<int64_t>(b, (int64_t)(I[k + 1] - I[k]));
                for (int q = I[k]; q < I[k + 1]; ++q)
                {
                    const int64_t g = J[q] < ncl ? B->first_col_diag + J[q] : B->col_map_offd[(size_t)(J[q] - ncl)];
                    put<int64_t>(b, g); put<double>(b, V[q]);
                }
            }
        }
        PE_TRY(exchange(hc, send, recv));
    }
    // 2. ghost columns G = B's own ghosts + those of the fetched rows
    std::vector<int64_t> &G = X.G;
    G.assign(B->col_map_offd.begin(), B->col_map_offd.end());
    std::vector<int32_t> eI(1, 0);
    std::vector<int64_t> eG; std::vector<double> eV;
    for (size_t r = 0; r < A->recv_procs.size(); ++r)
    {
        const std::vector<char> &b = recv[A->recv_procs[r]];
        const char *p = b.data(), *e = p + b.size();
        int rows = 0;
        while (p < e)
        {
            int64_t len; memcpy(&len, p, 8); p += 8;
            for (int64_t q = 0; q < len; ++q, p += 16) { int64_t g; double v; memcpy(&g, p, 8); memcpy(&v, p + 8, 8); eG.push_back(g); eV.push_back(v); }
            eI.push_back((int32_t)eG.size());
            ++rows;
        }
        PE_CHECK(rows == A->recv_vec_starts[r + 1] - A->recv_vec_starts[r], "distributed product: neighbour sent a wrong number of rows");
    }
    PE_CHECK((int)eI.size() - 1 == ngA, "distributed product: fetched rows do not cover the left factor's ghost columns");
    const int64_t c0 = X.c0, c1 = X.c1;
    for (int64_t g : eG) if (g < c0 || g >= c1) G.push_back(g);
    std::sort(G.begin(), G.end());
    G.erase(std::unique(G.begin(), G.end()), G.end());
    const int nce = X.nce();
    auto gpos = [&](int64_t g) { return ncl + (int)(std::lower_bound(G.begin(), G.end(), g) - G.begin()); };
    std::vector<int32_t> omap(B->col_map_offd.size()), eJ(eG.size());
    for (size_t j = 0; j < omap.size(); ++j) omap[j] = gpos(B->col_map_offd[j]);
    for (size_t q = 0; q < eG.size(); ++q) eJ[q] = (eG[q] >= c0 && eG[q] < c1) ? (int32_t)(eG[q] - c0) : gpos(eG[q]);
    // 3. B^ = [own rows ; fetched rows] on the device
    int *omap_d = nullptr;
    PE_CUDA(cudaMalloc(&omap_d, sizeof(int) * (omap.size() ? omap.size() : 1)));
    if (!omap.empty()) PE_CUDA(cudaMemcpyAsync(omap_d, omap.data(), sizeof(int) * omap.size(), cudaMemcpyHostToDevice, st));
    DevCSR Bloc, Bhat, Ahat;
    PE_TRY(merge_rows(ctx, B->diag, B->offd, nullptr, nB, omap_d, 0, nce, Bloc));
    PE_CUDA(cudaStreamSynchronize(st));
    cudaFree(omap_d);
    {
        const int64_t nnz_ext = (int64_t)eG.size();
        PE_TRY(devcsr_alloc(Bhat, nB + ngA, nce, Bloc.nnz + nnz_ext));
        PE_CUDA(cudaMemcpyAsync(Bhat.I, Bloc.I, sizeof(int) * (size_t)(nB + 1), cudaMemcpyDeviceToDevice, st));
        if (Bloc.nnz)
        {
            PE_CUDA(cudaMemcpyAsync(Bhat.J, Bloc.J, sizeof(int) * (size_t)Bloc.nnz, cudaMemcpyDeviceToDevice, st));
            PE_CUDA(cudaMemcpyAsync(Bhat.A, Bloc.A, sizeof(double) * (size_t)Bloc.nnz, cudaMemcpyDeviceToDevice, st));
        }
        std::vector<int32_t> eIs(eI.begin() + 1, eI.end());
        for (auto &v : eIs) v += (int32_t)Bloc.nnz;
        if (ngA) PE_CUDA(cudaMemcpyAsync(Bhat.I + nB + 1, eIs.data(), sizeof(int) * (size_t)ngA, cudaMemcpyHostToDevice, st));
        if (nnz_ext)
        {
            PE_CUDA(cudaMemcpyAsync(Bhat.J + Bloc.nnz, eJ.data(), sizeof(int) * (size_t)nnz_ext, cudaMemcpyHostToDevice, st));
            PE_CUDA(cudaMemcpyAsync(Bhat.A + Bloc.nnz, eV.data(), sizeof(double) * (size_t)nnz_ext, cudaMemcpyHostToDevice, st));
        }
        PE_CUDA(cudaStreamSynchronize(st));
    }
    devcsr_free(Bloc);
    // 4. A^ = [A_diag | A_offd];  AB = A^ B^
    PE_TRY(merge_rows(ctx, A->diag, A->offd, nullptr, nA, nullptr, nB, nB + ngA, Ahat));
    PE_TRY(pe_devcsr_spgemm(ctx, Ahat, Bhat, X.AB));
    devcsr_free(Ahat); devcsr_free(Bhat);
    return 0;
}

// local CSR on extended index spaces -> ParCSR (rows: global ids / owners; mode 0 = contributions to rows of other ranks
// are sent to their owners and summed, 1 = all rows are mine)
static int finish_parcsr(pe_ctx *ctx, DevCSR &C, int mode, const std::vector<int64_t> &rgid, const std::vector<int32_t> &rown,
                         const std::vector<int64_t> &cgid, const std::vector<int32_t> &cown, int64_t r0, int64_t r1, int64_t rglob,
                         int64_t c0, int64_t c1, int64_t cglob, pe_mat **out)
{
    std::vector<int32_t> CI, CJ; std::vector<double> CA;
    PE_TRY(download_csr(ctx, C, CI, CJ, CA));
    devcsr_free(C);
    pe_parcsr_owned *M = nullptr;
    PE_TRY(pe_par_assemble(ctx->hcomm, mode, (int32_t)rgid.size(), (int32_t)cgid.size(), CI.data(), CJ.data(), CA.data(), rgid.data(), rown.data(),
                           cgid.data(), cown.data(), r0, r1, rglob, c0, c1, cglob, &M));
    const int rc = pe_mat_upload(ctx, pe_parcsr_owned_view(M), out);
    pe_parcsr_owned_free(M);
    return rc;
}
static void own_rows(const pe_mat *A, int me, std::vector<int64_t> &gid, std::vector<int32_t> &own)
{
    gid.resize((size_t)A->diag.nrows); own.assign((size_t)A->diag.nrows, me);
    for (int i = 0; i < A->diag.nrows; ++i) gid[i] = A->first_row_index + i;
}

// ---------------------------------------------------------------------------------------------
// C = A B for distributed operands (hypre_ParMatmul behind mfem::ParMult, SchurComplementFactory.cpp:133)
// ---------------------------------------------------------------------------------------------
int pe_spgemm_distributed(pe_ctx *ctx, const pe_mat *A, const pe_mat *B, pe_mat **C)
{
    ExtProduct X;
    PE_TRY(ext_product(ctx, A, B, X));
    std::vector<int64_t> rgid, cgid; std::vector<int32_t> rown, cown;
    own_rows(A, ctx->hcomm->rank, rgid, rown);
    X.col_ids(ctx->hcomm->rank, cgid, cown);
    return finish_parcsr(ctx, X.AB, 1, rgid, rown, cgid, cown, A->first_row_index, A->first_row_index + A->diag.nrows, A->global_num_rows,
                         X.c0, X.c1, B->global_num_cols, C);
}

// ---------------------------------------------------------------------------------------------
// distributed R^T A P (hypre_BoomerAMGBuildCoarseOperator behind mfem::RAP; R = P: Hierarchy.cpp:365, R != P: the
// off-diagonal blocks of the blocked hierarchy, Hierarchy.cpp:400-544):
//   C^ = (R^ own rows)^T (A^ P^), R^ = [R_diag | R_offd]; rows of C^ that belong to ghost coarse dofs of R are
//   contributions to other ranks' rows: Assemble sends them to their owners and sums
// ---------------------------------------------------------------------------------------------
int pe_rap_distributed(pe_ctx *ctx, const pe_mat *R, const pe_mat *A, const pe_mat *P, pe_mat **Ac)
{
    const pe_host_comm *hc = ctx->hcomm;
    PE_CHECK(hc, "distributed RAP needs a host communicator (pe_ctx_set_host_comm)");
    if (!R) R = P;
    PE_CHECK(R->diag.nrows == A->diag.nrows && A->diag.ncols == P->diag.nrows, "pe_rap: size mismatch");
    ExtProduct X;
    PE_TRY(ext_product(ctx, A, P, X));
    const int ncr = R->diag.ncols, ngr = R->offd.ncols;
    DevCSR Rloc, RlocT, C;
    PE_TRY(merge_rows(ctx, R->diag, R->offd, nullptr, R->diag.nrows, nullptr, ncr, ncr + ngr, Rloc));
    PE_TRY(pe_devcsr_transpose(ctx, Rloc, RlocT));
    devcsr_free(Rloc);
    PE_TRY(pe_devcsr_spgemm(ctx, RlocT, X.AB, C));
    devcsr_free(RlocT); devcsr_free(X.AB);
    std::vector<int64_t> rstart;
    PE_TRY(col_partition(hc, R, rstart));
    std::vector<int64_t> rgid((size_t)ncr + ngr), cgid;
    std::vector<int32_t> rown((size_t)ncr + ngr, hc->rank), cown;
    for (int j = 0; j < ncr; ++j) rgid[j] = R->first_col_diag + j;
    for (int j = 0; j < ngr; ++j)
    {
        rgid[ncr + j] = R->col_map_offd[(size_t)j];
        rown[ncr + j] = (int32_t)(std::upper_bound(rstart.begin(), rstart.end(), rgid[ncr + j]) - rstart.begin()) - 1;
    }
    X.col_ids(hc->rank, cgid, cown);
    return finish_parcsr(ctx, C, 0, rgid, rown, cgid, cown, R->first_col_diag, R->first_col_diag + ncr, R->global_num_cols,
                         X.c0, X.c1, P->global_num_cols, Ac);
}

// ---------------------------------------------------------------------------------------------
// A^T for a distributed A (hypre_ParCSRMatrixTranspose): the transposed ghost columns are rows of other ranks
// ---------------------------------------------------------------------------------------------
int pe_transpose_distributed(pe_ctx *ctx, const pe_mat *A, pe_mat **out)
{
    const pe_host_comm *hc = ctx->hcomm;
    PE_CHECK(hc, "distributed transpose needs a host communicator (pe_ctx_set_host_comm)");
    const int nc = A->diag.ncols, ng = A->offd.ncols;
    DevCSR Aloc, T;
    PE_TRY(merge_rows(ctx, A->diag, A->offd, nullptr, A->diag.nrows, nullptr, nc, nc + ng, Aloc));
    PE_TRY(pe_devcsr_transpose(ctx, Aloc, T));
    devcsr_free(Aloc);
    std::vector<int64_t> cstart;
    PE_TRY(col_partition(hc, A, cstart));
    std::vector<int64_t> rgid((size_t)nc + ng), cgid; std::vector<int32_t> rown((size_t)nc + ng, hc->rank), cown;
    for (int j = 0; j < nc; ++j) rgid[j] = A->first_col_diag + j;
    for (int j = 0; j < ng; ++j)
    {
        rgid[nc + j] = A->col_map_offd[(size_t)j];
        rown[nc + j] = (int32_t)(std::upper_bound(cstart.begin(), cstart.end(), rgid[nc + j]) - cstart.begin()) - 1;
    }
    own_rows(A, hc->rank, cgid, cown);
    return finish_parcsr(ctx, T, 0, rgid, rown, cgid, cown, A->first_col_diag, A->first_col_diag + nc, A->global_num_cols,
                         A->first_row_index, A->first_row_index + A->diag.nrows, A->global_num_rows, out);
}

// ---------------------------------------------------------------------------------------------
// C = a A + b B for distributed operands with the same row and column partition (hypre_ParCSRMatrixAdd2,
// src/hypreExtension/parcsr-add.c: merge diag and offd, add, split again)
// ---------------------------------------------------------------------------------------------
int pe_devcsr_spadd(pe_ctx *ctx, double a, const DevCSR &A, double b, const DevCSR &B, DevCSR &C);   // pe_spgemm.cu
int pe_spadd_distributed(pe_ctx *ctx, double a, const pe_mat *A, double b, const pe_mat *B, pe_mat **out)
{
    const pe_host_comm *hc = ctx->hcomm;
    PE_CHECK(hc, "distributed matrix addition needs a host communicator (pe_ctx_set_host_comm)");
    PE_CHECK(A->diag.nrows == B->diag.nrows && A->diag.ncols == B->diag.ncols && A->first_col_diag == B->first_col_diag &&
             A->global_num_cols == B->global_num_cols, "pe_spadd: operands are partitioned differently");
    cudaStream_t st = ctx->stream;
    const int n = A->diag.nrows, nc = A->diag.ncols;
    std::vector<int64_t> U(A->col_map_offd.begin(), A->col_map_offd.end());
    U.insert(U.end(), B->col_map_offd.begin(), B->col_map_offd.end());
    std::sort(U.begin(), U.end());
    U.erase(std::unique(U.begin(), U.end()), U.end());
    auto remap = [&](const pe_mat *M, DevCSR &L) -> int
    {
        std::vector<int32_t> om(M->col_map_offd.size());
        for (size_t j = 0; j < om.size(); ++j) om[j] = nc + (int32_t)(std::lower_bound(U.begin(), U.end(), M->col_map_offd[j]) - U.begin());
        int *om_d = nullptr;
        PE_CUDA(cudaMalloc(&om_d, sizeof(int) * (om.size() ? om.size() : 1)));
        if (!om.empty()) PE_CUDA(cudaMemcpyAsync(om_d, om.data(), sizeof(int) * om.size(), cudaMemcpyHostToDevice, st));
        const int rc = merge_rows(ctx, M->diag, M->offd, nullptr, n, om_d, 0, nc + (int)U.size(), L);
        cudaStreamSynchronize(st);
        cudaFree(om_d);
        return rc;
    };
    DevCSR LA, LB, C;
    PE_TRY(remap(A, LA));
    PE_TRY(remap(B, LB));
    PE_TRY(pe_devcsr_spadd(ctx, a, LA, b, LB, C));
    devcsr_free(LA); devcsr_free(LB);
    std::vector<int64_t> cstart;
    PE_TRY(col_partition(hc, A, cstart));
    std::vector<int64_t> rgid, cgid((size_t)nc + U.size()); std::vector<int32_t> rown, cown((size_t)nc + U.size(), hc->rank);
    own_rows(A, hc->rank, rgid, rown);
    for (int j = 0; j < nc; ++j) cgid[j] = A->first_col_diag + j;
    for (size_t j = 0; j < U.size(); ++j)
    {
        cgid[nc + j] = U[j];
        cown[nc + j] = (int32_t)(std::upper_bound(cstart.begin(), cstart.end(), U[j]) - cstart.begin()) - 1;
    }
    return finish_parcsr(ctx, C, 1, rgid, rown, cgid, cown, A->first_row_index, A->first_row_index + n, A->global_num_rows,
                         A->first_col_diag, A->first_col_diag + nc, A->global_num_cols, out);
}

// ---------------------------------------------------------------------------------------------
// y = alpha A^T x + beta y for a distributed A (hypre_ParCSRMatrixMatvecT):
//   y_ext = offd^T x -> reverse exchange to the owners -> y = alpha diag^T x + beta y -> owners add
//   the received partial sums in a fixed order (deterministic gather, no atomics)
// ---------------------------------------------------------------------------------------------
__global__ void k_unpack_add(int nt, const int *__restrict__ rows, const int *__restrict__ uI, const int *__restrict__ upos,
                             const double *__restrict__ buf, double alpha, double *y)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    double s = 0.0;
    for (int k = uI[t]; k < uI[t + 1]; ++k) s += buf[upos[k]];
    y[rows[t]] += alpha * s;
}

int pe_reverse_halo_add(pe_mat *A, double alpha, double *y_d);   // pe_core.cu (needs the NCCL table)

int pe_spmv_t_distributed(pe_ctx *ctx, double alpha, pe_mat *A, const pe_vec *x, double beta, pe_vec *y)
{
    if (!A->T)
    {
        DevCSR T;
        PE_TRY(pe_devcsr_transpose(ctx, A->diag, T));
        PE_TRY(pe_mat_wrap_local(ctx, T, &A->T));
    }
    if (A->offd.ncols > 0 && !A->offdT_built)
    {
        PE_TRY(pe_devcsr_transpose(ctx, A->offd, A->offdT));
        A->offdT_built = true;
    }
    if (!A->unpack_built)
    {
        // distinct target rows of the reverse exchange and, per row, the buffer slots that feed it
        const int nsend = A->send_map_starts.empty() ? 0 : A->send_map_starts.back();
        std::vector<std::pair<int, int>> v((size_t)nsend);
        for (int k = 0; k < nsend; ++k) v[k] = {A->send_map_elmts[k], k};
        std::sort(v.begin(), v.end());
        std::vector<int> rows, uI(1, 0), upos;
        for (int k = 0; k < nsend; ++k)
        {
            if (k == 0 || v[k].first != v[k - 1].first) { if (k) uI.push_back(k); rows.push_back(v[k].first); }
            upos.push_back(v[k].second);
        }
        if (nsend) uI.push_back(nsend);
        A->n_unpack = (int)rows.size();
        if (A->n_unpack)
        {
            PE_CUDA(cudaMalloc(&A->unpack_rows_d, sizeof(int) * rows.size()));
            PE_CUDA(cudaMalloc(&A->unpack_I_d, sizeof(int) * uI.size()));
            PE_CUDA(cudaMalloc(&A->unpack_pos_d, sizeof(int) * upos.size()));
            PE_CUDA(cudaMemcpy(A->unpack_rows_d, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice));
            PE_CUDA(cudaMemcpy(A->unpack_I_d, uI.data(), sizeof(int) * uI.size(), cudaMemcpyHostToDevice));
            PE_CUDA(cudaMemcpy(A->unpack_pos_d, upos.data(), sizeof(int) * upos.size(), cudaMemcpyHostToDevice));
        }
        A->unpack_built = true;
    }
    // partial sums for the ghost columns, then the local part, then the owners' additions
    if (A->offd.ncols > 0)
        PE_TRY(pe_launch_spmv(ctx, A->offdT, nullptr, pe_choose_tpr(A->offdT.nnz, A->offdT.nrows), 1.0, x->d, nullptr, 0.0, A->x_ext_d, A->x_ext_d));
    PE_TRY(pe_launch_spmv(ctx, A->T->diag, nullptr, A->T->tpr, alpha, x->d, nullptr, beta, y->d, y->d));
    PE_TRY(pe_reverse_halo_add(A, alpha, y->d));
    return 0;
}

int pe_launch_unpack_add(pe_ctx *ctx, pe_mat *A, double alpha, double *y_d)
{
    if (A->n_unpack == 0) return 0;
    k_unpack_add<<<pe_grid_for(A->n_unpack, 256), 256, 0, ctx->stream>>>(A->n_unpack, A->unpack_rows_d, A->unpack_I_d, A->unpack_pos_d,
                                                                        A->send_buf_d, alpha, y_d);
    PE_LAUNCHED(ctx);
    return 0;
}
