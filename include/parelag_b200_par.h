/*
 * parelag_b200_par.h -- multi-rank (one rank <-> one mesh partition <-> one GPU) part of
 * the C ABI: the host communicator the setup phase uses, and the entry points that
 * replace ParElag's SharingMap / hypre ParCSR assembly on the path.
 *
 * Data-path communication (ParCSR halo exchange in SpMV / relaxation, dot products) is
 * NCCL on the device (pe_ctx_create with a unique id, parelag_b200.h).  SETUP-time
 * exchanges of host integer tables and matrix rows (the MPI calls inside hypre's
 * ParCSR RAP / MatvecCommPkgCreate and ParElag's SharedEntityCommunication) go through
 * the callback table below, which the embedding application fills: MPI in a ParElag
 * build, torch.distributed(gloo) in this repository's harness.
 */
#ifndef PARELAG_B200_PAR_H
#define PARELAG_B200_PAR_H
#include "parelag_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pe_host_comm {
    int rank, size;
    void *user;
    /* every rank contributes nbytes; recv holds size*nbytes, rank-major (MPI_Allgather) */
    int (*allgather)(void *user, const void *send, int64_t nbytes, void *recv);
    /* MPI_Alltoallv on bytes: send_bytes[r] bytes at send+send_displs[r] go to rank r */
    int (*alltoallv)(void *user, const void *send, const int64_t *send_bytes, const int64_t *send_displs,
                     void *recv, const int64_t *recv_bytes, const int64_t *recv_displs);
} pe_host_comm;

/* attach the host communicator to a context (borrowed; must outlive the context's use) */
int pe_ctx_set_host_comm(pe_ctx *ctx, const pe_host_comm *comm);

/* ---- SharingMap::Assemble(range, A, domain) = E_r^T A E_d and IgnoreNonLocalRange
 * (src/structures/SharingMap.cpp:975-1011, 930-946) for a rank-local CSR matrix in local
 * dof numbering.  row_gid / col_gid: global TRUE id of every local row / column;
 * row_owner / col_owner: owning rank.  mode 0 = Assemble (contributions of all ranks to a
 * true row are summed on its owner), 1 = IgnoreNonLocalRange (rows the caller owns only).
 * The result is a host ParCSR with hypre's layout (diag / offd / ascending col_map_offd /
 * comm package), owned by the handle; view() is valid until free(). */
typedef struct pe_parcsr_owned pe_parcsr_owned;
int pe_par_assemble(const pe_host_comm *comm, int mode, int32_t nrows, int32_t ncols, const int32_t *I,
                    const int32_t *J, const double *A, const int64_t *row_gid, const int32_t *row_owner,
                    const int64_t *col_gid, const int32_t *col_owner, int64_t row_start, int64_t row_end,
                    int64_t global_rows, int64_t col_start, int64_t col_end, int64_t global_cols,
                    pe_parcsr_owned **out);
const pe_parcsr_host *pe_parcsr_owned_view(const pe_parcsr_owned *M);
int pe_parcsr_owned_free(pe_parcsr_owned *M);

/* hypre_MatvecCommPkgCreate: (re)build the comm package of a host ParCSR whose diag / offd /
 * col_map_offd are set; col_starts[size+1] = first owned column of every rank */
int pe_par_build_comm_pkg(const pe_host_comm *comm, pe_parcsr_owned *M, const int64_t *col_starts);

/* ---- true-entity numbering (SharingMap::SetUp for shared entities, SharingMap.cpp:213-497):
 * every local item carries a global key (identical on all ranks that hold a copy) and the
 * list of ranks holding it (CSR sharers_I / sharers_J, own rank included; a private item may
 * have an empty list).  Owner = smallest rank.  Output: global true id and owner of every
 * local item; returns the first global id owned by this rank and the global count. */
int pe_par_number_items(const pe_host_comm *comm, int32_t n, const int64_t *key, const int32_t *sharers_I,
                        const int32_t *sharers_J, int64_t *gid, int32_t *owner, int64_t *my_start,
                        int64_t *my_count, int64_t *global_count);

#ifdef __cplusplus
}
#endif
#endif /* PARELAG_B200_PAR_H */
