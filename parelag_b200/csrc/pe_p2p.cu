// pe_p2p.cu -- ParCSR halo exchange over NVLink peer memory (one node, one process per GPU).
//
// Replaces the pack kernel + grouped ncclSend/ncclRecv of hypre_ParCSRMatrixMatvec's comm
// handle (hypre_ParCSRMatrixMatvecBoolInt.c:162-199) by ONE kernel per exchange that gathers
// x[send_map_elmts] and stores the values straight into the neighbours' ghost buffers over
// NVLink, plus a one-CTA kernel that waits for the neighbours' arrival flags.  No host call, no
// NCCL launch: the exchange is two ordinary kernel nodes of the V-cycle graph.
//
// Memory: every rank owns one "halo arena" (cudaMalloc, exported with cudaIpcGetMemHandle and
// opened by all other ranks at pe_ctx_set_host_comm).  Ghost buffers (x_ext), reverse-exchange
// receive buffers (send_buf) and all flags of every distributed matrix are carved out of it, so a
// rank addresses its neighbours' buffers as peer_base[rank] + offset; the offsets are exchanged
// once per matrix through the host communicator (pe_p2p_link, collective, at the first exchange).
//
// Protocol per matrix and direction, exchange number e = 1, 2, ... (seq lives in device memory, so a
// captured graph replays correctly):
//   k_halo_push : (a) tells every rank I receive from that I have consumed exchange e-1 (ack = e-1;
//                 stream order guarantees my consumers of e-1 are complete), (b) waits until the
//                 ranks I send to have acknowledged e-1 (their buffer is free), (c) stores my
//                 values into their buffers, (d) the last CTA publishes arrival flag = e to them
//                 (system-scope release).
//   k_halo_wait : spins until every rank I receive from has published e (system-scope acquire),
//                 then sets seq = e.  Consumers follow in stream order.
// All ranks execute the same sequence of exchanges (SPMD), so (a) is always reached without
// depending on data of exchange e: no deadlock.  Forward (ghost fill) and reverse (MatvecT partial
// sums back to the owners) use separate flags and counters.
#include "pe_core.cuh"
#include "../../include/parelag_b200_par.h"
#include <cstring>

typedef unsigned long long u64;

struct PeP2PDir {
    int nout = 0, nin = 0, ntot = 0;
    int *out_starts_d = nullptr;     // nout+1: element ranges per destination rank
    double **out_dst_d = nullptr;    // nout: remote buffer segment of every destination
    u64 **out_flag_d = nullptr;      // nout: remote arrival flag I publish to
    u64 *ack_local_d = nullptr;      // nout: acknowledgements of my destinations (in my arena)
    u64 *flag_local_d = nullptr;     // nin : arrival flags of my sources (in my arena)
    u64 **ack_remote_d = nullptr;    // nin : remote ack slot of every source
    u64 *seq_d = nullptr;            // completed exchanges
    unsigned int *done_d = nullptr;  // CTA arrival counter of the push kernel
};

__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p)
{
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256)
k_halo_push(PeP2PDir d, const int32_t *__restrict__ map, const double *__restrict__ src)
{
    const u64 e = *((volatile u64 *)d.seq_d) + 1;
    // (a) my consumers of exchange e-1 are done: release the sources' send slots
    if (blockIdx.x == 0 && threadIdx.x < d.nin) st_release_sys(d.ack_remote_d[threadIdx.x], e - 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.ntot)
    {
        int nb = 0;
        while (nb + 1 < d.nout && i >= d.out_starts_d[nb + 1]) ++nb;
        // (b) destination has consumed exchange e-1
        while (ld_acquire_sys(d.ack_local_d + nb) < e - 1) { }
        // (c) gather and store over NVLink
        d.out_dst_d[nb][i - d.out_starts_d[nb]] = map ? src[map[i]] : src[i];
    }
    // (d) last CTA publishes the arrival flags
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(d.done_d, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last)
    {
        __threadfence_system();
        if (threadIdx.x < d.nout) st_release_sys(d.out_flag_d[threadIdx.x], e);
        if (threadIdx.x == 0) *d.done_d = 0u;
    }
}

__global__ void k_halo_wait(PeP2PDir d)
{
    const u64 e = *((volatile u64 *)d.seq_d) + 1;
    if (threadIdx.x < d.nin) while (ld_acquire_sys(d.flag_local_d + threadIdx.x) < e) { }
    __syncthreads();
    if (threadIdx.x == 0) { *d.seq_d = e; __threadfence(); }
}

// ---------------------------------------------------------------------------------------------
// arena
// ---------------------------------------------------------------------------------------------
int pe_p2p_init(pe_ctx *ctx)
{
    if (ctx->nranks <= 1 || !ctx->hcomm || ctx->p2p_base || pe_get_tuning(PE_TUNE_P2P_HALO) == 0) return 0;
    const pe_host_comm *hc = ctx->hcomm;
    const size_t bytes = (size_t)(getenv("PE_P2P_ARENA_MB") ? atoll(getenv("PE_P2P_ARENA_MB")) : 256) << 20;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    char *base = nullptr;
    int ok = cudaMalloc(&base, bytes) == cudaSuccess ? 1 : 0;
    if (ok) ok = cudaMemset(base, 0, bytes) == cudaSuccess && cudaIpcGetMemHandle(&mine, base) == cudaSuccess ? 1 : 0;
    if (!ok) (void)cudaGetLastError();
    struct Rec { cudaIpcMemHandle_t h; int ok; int pad; };
    Rec me;
    memset(&me, 0, sizeof me);
    me.h = mine; me.ok = ok;
    std::vector<Rec> all((size_t)hc->size);
    PE_CHECK(hc->allgather(hc->user, &me, (int64_t)sizeof(Rec), all.data()) == 0, "pe_p2p_init: host allgather failed");
    int all_ok = 1;
    for (int r = 0; r < hc->size; ++r) all_ok &= all[r].ok;
    std::vector<char *> peer((size_t)hc->size, nullptr);
    if (all_ok)
        for (int r = 0; r < hc->size && ok; ++r)
        {
            if (r == ctx->rank) { peer[r] = base; continue; }
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; (void)cudaGetLastError(); }
            peer[r] = static_cast<char *>(p);
        }
    // every rank must have mapped every arena, or nobody uses the peer path
    int mapped = all_ok && ok ? 1 : 0;
    std::vector<int> votes((size_t)hc->size, 0);
    PE_CHECK(hc->allgather(hc->user, &mapped, (int64_t)sizeof(int), votes.data()) == 0, "pe_p2p_init: host allgather failed");
    for (int r = 0; r < hc->size; ++r) mapped &= votes[r];
    if (!mapped)
    {
        for (int r = 0; r < hc->size; ++r) if (r != ctx->rank && peer[r]) cudaIpcCloseMemHandle(peer[r]);
        if (base) cudaFree(base);
        (void)cudaGetLastError();
        return 0;                       // the NCCL path stays in use
    }
    ctx->p2p_base = base; ctx->p2p_size = bytes; ctx->p2p_used = 0; ctx->p2p_peer = peer;
    return 0;
}

void pe_p2p_shutdown(pe_ctx *ctx)
{
    if (!ctx->p2p_base) return;
    for (size_t r = 0; r < ctx->p2p_peer.size(); ++r)
        if ((int)r != ctx->rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    cudaFree(ctx->p2p_base);
    ctx->p2p_base = nullptr;
    ctx->p2p_peer.clear();
}

// bump allocation (256-byte granules); null when the arena is absent or full -> caller uses cudaMalloc / NCCL
void *pe_p2p_alloc(pe_ctx *ctx, size_t bytes)
{
    if (!ctx->p2p_base) return nullptr;
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (ctx->p2p_used + need > ctx->p2p_size) return nullptr;
    void *p = ctx->p2p_base + ctx->p2p_used;
    ctx->p2p_used += need;
    return p;
}
bool pe_p2p_owns(const pe_ctx *ctx, const void *p)
{
    return ctx->p2p_base && (const char *)p >= ctx->p2p_base && (const char *)p < ctx->p2p_base + ctx->p2p_size;
}

// ---------------------------------------------------------------------------------------------
// link: exchange buffer / flag offsets with the neighbours (collective)
// ---------------------------------------------------------------------------------------------
static void dir_free(PeP2PDir *d)
{
    if (!d) return;
    cudaFree(d->out_starts_d); cudaFree(d->out_dst_d); cudaFree(d->out_flag_d); cudaFree(d->ack_remote_d);
    delete d;
}
void pe_p2p_unlink(pe_mat *A)
{
    dir_free(A->p2p[0]); dir_free(A->p2p[1]);
    A->p2p[0] = A->p2p[1] = nullptr;
}

template <class T> static int up_array(const std::vector<T> &h, T **d)
{
    PE_CUDA(cudaMalloc(d, sizeof(T) * (h.empty() ? 1 : h.size())));
    if (!h.empty()) PE_CUDA(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}

int pe_p2p_link(pe_mat *A)
{
    pe_ctx *ctx = A->ctx;
    const pe_host_comm *hc = ctx->hcomm;
    A->p2p_state = -1;
    const int ns = (int)A->send_procs.size(), nr = (int)A->recv_procs.size();
    const int nsend = ns ? A->send_map_starts.back() : 0, nrecv = nr ? A->recv_vec_starts.back() : 0;
    // local resources; `mine` = 0 if this rank cannot take part (buffers outside the arena, arena full)
    int mine = ctx->p2p_base ? 1 : 0;
    if (nrecv > 0 && !pe_p2p_owns(ctx, A->x_ext_d)) mine = 0;
    if (nsend > 0 && !pe_p2p_owns(ctx, A->send_buf_d)) mine = 0;
    u64 *flags = nullptr;          // [fwd ack ns | fwd flag nr | rev ack nr | rev flag ns | seq 2 | done 2]
    const size_t nflags = (size_t)2 * ns + 2 * nr + 4;
    if (mine) { flags = static_cast<u64 *>(pe_p2p_alloc(ctx, sizeof(u64) * nflags)); if (!flags) mine = 0; }
    // table row for rank t: {fwd data off, fwd flag off, fwd ack off, rev data off, rev flag off, rev ack off, ok}
    const int W = 7;
    std::vector<int64_t> row((size_t)hc->size * W, -1), all((size_t)hc->size * hc->size * W, -1);
    for (int t = 0; t < hc->size; ++t) row[(size_t)t * W + 6] = mine;
    if (mine)
    {
        const char *base = ctx->p2p_base;
        u64 *fwd_ack = flags, *fwd_flag = flags + ns, *rev_ack = flags + ns + nr, *rev_flag = flags + ns + 2 * nr;
        for (int r = 0; r < nr; ++r)
        {
            int64_t *e = &row[(size_t)A->recv_procs[r] * W];
            e[0] = (const char *)(A->x_ext_d + A->recv_vec_starts[r]) - base;     // where recv_procs[r] stores my ghosts
            e[1] = (const char *)(fwd_flag + r) - base;
            e[5] = (const char *)(rev_ack + r) - base;                            // it acknowledges my reverse sends here
        }
        for (int s = 0; s < ns; ++s)
        {
            int64_t *e = &row[(size_t)A->send_procs[s] * W];
            e[2] = (const char *)(fwd_ack + s) - base;                            // it acknowledges my forward sends here
            e[3] = (const char *)(A->send_buf_d + A->send_map_starts[s]) - base;  // where it stores partial sums for me
            e[4] = (const char *)(rev_flag + s) - base;
        }
    }
    PE_CHECK(hc->allgather(hc->user, row.data(), (int64_t)(sizeof(int64_t) * row.size()), all.data()) == 0,
             "pe_p2p_link: host allgather failed");
    for (int t = 0; t < hc->size; ++t)
        if (all[((size_t)t * hc->size + ctx->rank) * W + 6] != 1) return 0;       // some rank opted out: everybody keeps NCCL
    auto told = [&](int t, int k) { return all[((size_t)t * hc->size + ctx->rank) * W + k]; };
    u64 *fwd_ack = flags, *fwd_flag = flags + ns, *rev_ack = flags + ns + nr, *rev_flag = flags + ns + 2 * nr;
    u64 *seq = flags + 2 * ns + 2 * nr;
    unsigned int *done = reinterpret_cast<unsigned int *>(seq + 2);
    // From here on a failure of ONE rank (asymmetric comm package, allocation) must not leave the others on the
    // peer path spinning on flags nobody publishes: build locally, then vote (as pe_p2p_init does).
    auto build = [&]() -> int {
    for (int dir = 0; dir < 2; ++dir)
    {
        PeP2PDir *d = new PeP2PDir();
        A->p2p[dir] = d;
        const std::vector<int32_t> &outp = dir == 0 ? A->send_procs : A->recv_procs;
        const std::vector<int32_t> &inp = dir == 0 ? A->recv_procs : A->send_procs;
        const std::vector<int32_t> &starts = dir == 0 ? A->send_map_starts : A->recv_vec_starts;
        d->nout = (int)outp.size(); d->nin = (int)inp.size(); d->ntot = d->nout ? starts.back() : 0;
        PE_CHECK(d->nout <= 256 && d->nin <= 256, "pe_p2p_link: more than 256 neighbours");
        std::vector<int> st(starts.begin(), starts.end());
        if (st.empty()) st.push_back(0);
        std::vector<double *> dst((size_t)d->nout);
        std::vector<u64 *> oflag((size_t)d->nout), rack((size_t)d->nin);
        for (int k = 0; k < d->nout; ++k)
        {
            const int t = outp[k];
            const int64_t doff = told(t, dir == 0 ? 0 : 3), foff = told(t, dir == 0 ? 1 : 4);
            PE_CHECK(doff >= 0 && foff >= 0, "pe_p2p_link: neighbour does not list this rank (asymmetric comm package)");
            dst[k] = reinterpret_cast<double *>(ctx->p2p_peer[t] + doff);
            oflag[k] = reinterpret_cast<u64 *>(ctx->p2p_peer[t] + foff);
        }
        for (int k = 0; k < d->nin; ++k)
        {
            const int t = inp[k];
            const int64_t aoff = told(t, dir == 0 ? 2 : 5);
            PE_CHECK(aoff >= 0, "pe_p2p_link: neighbour does not list this rank (asymmetric comm package)");
            rack[k] = reinterpret_cast<u64 *>(ctx->p2p_peer[t] + aoff);
        }
        PE_TRY(up_array(st, &d->out_starts_d));
        PE_TRY(up_array(dst, &d->out_dst_d));
        PE_TRY(up_array(oflag, &d->out_flag_d));
        PE_TRY(up_array(rack, &d->ack_remote_d));
        d->ack_local_d = dir == 0 ? fwd_ack : rev_ack;
        d->flag_local_d = dir == 0 ? fwd_flag : rev_flag;
        d->seq_d = seq + dir;
        d->done_d = done + dir;
    }
    return 0;
    };
    int built = build() == 0 ? 1 : 0;
    std::vector<int> votes((size_t)hc->size, 0);
    PE_CHECK(hc->allgather(hc->user, &built, (int64_t)sizeof(int), votes.data()) == 0, "pe_p2p_link: host allgather failed");
    for (int r = 0; r < hc->size; ++r) built &= votes[r];
    if (!built) { pe_p2p_unlink(A); return 0; }       // everybody keeps NCCL for this matrix (state stays -1)
    A->p2p_state = 1;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// exchange
// ---------------------------------------------------------------------------------------------
int pe_p2p_push(pe_mat *A, int dir, const double *src)
{
    pe_ctx *c = A->ctx;
    const PeP2PDir &d = *A->p2p[dir];
    if (d.nout == 0 && d.nin == 0) return 0;
    k_halo_push<<<pe_grid_for(d.ntot, 256), 256, 0, c->stream>>>(d, dir == 0 ? A->send_map_d : nullptr, src);
    PE_LAUNCHED(c);
    return 0;
}
int pe_p2p_wait(pe_mat *A, int dir)
{
    pe_ctx *c = A->ctx;
    const PeP2PDir &d = *A->p2p[dir];
    if (d.nout == 0 && d.nin == 0) return 0;
    k_halo_wait<<<1, 256, 0, c->stream>>>(d);
    PE_LAUNCHED(c);
    return 0;
}
