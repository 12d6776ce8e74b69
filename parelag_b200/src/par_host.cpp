// par_host.cpp -- extern "C" entry points of include/parelag_b200_par.h that run on the host:
// true-item numbering, Assemble / IgnoreNonLocalRange into hypre's ParCSR layout, comm package.
#include "par_host.hpp"
#include <memory>
#include <numeric>

void pe_set_error(const std::string &msg);   // pe_core.cu

struct pe_parcsr_owned
{
    pe_parcsr_host H{};
    std::vector<int32_t> diag_i, diag_j, offd_i, offd_j;
    std::vector<double> diag_a, offd_a;
    std::vector<int64_t> col_map_offd;
    std::vector<int32_t> send_procs, send_map_starts, send_map_elmts, recv_procs, recv_vec_starts;
    void refresh()
    {
        H.diag_i = diag_i.data(); H.diag_j = diag_j.data(); H.diag_data = diag_a.data();
        H.offd_i = offd_i.data(); H.offd_j = offd_j.data(); H.offd_data = offd_a.data();
        H.num_cols_offd = (int32_t)col_map_offd.size();
        H.col_map_offd = col_map_offd.data();
        H.num_sends = (int32_t)send_procs.size(); H.send_procs = send_procs.data();
        H.send_map_starts = send_map_starts.data(); H.send_map_elmts = send_map_elmts.data();
        H.num_recvs = (int32_t)recv_procs.size(); H.recv_procs = recv_procs.data();
        H.recv_vec_starts = recv_vec_starts.data();
    }
};

#define PAR_TRY try {
#define PAR_CATCH                                                 \
    }                                                             \
    catch (const std::exception &e) { pe_set_error(e.what()); return 5; } \
    return 0;

using namespace parelag::par;

extern "C" int pe_par_number_items(const pe_host_comm *comm, int32_t n, const int64_t *key, const int32_t *sI,
                                   const int32_t *sJ, int64_t *gid, int32_t *owner, int64_t *my_start,
                                   int64_t *my_count, int64_t *global_count)
{
    PAR_TRY
    const int me = comm->rank, np = comm->size;
    int64_t mine = 0;
    for (int i = 0; i < n; ++i)
    {
        int o = me;
        for (int k = sI[i]; k < sI[i + 1]; ++k) o = std::min(o, (int)sJ[k]);
        owner[i] = o;
        if (o == me) ++mine;
    }
    std::vector<int64_t> counts = AllGather<int64_t>(comm, mine);
    int64_t start = 0, total = 0;
    for (int r = 0; r < np; ++r) { if (r < me) start += counts[r]; total += counts[r]; }
    int64_t next = start;
    for (int i = 0; i < n; ++i) gid[i] = owner[i] == me ? next++ : -1;
    // owners tell the other holders the global id, keyed by the item's global key
    std::vector<std::vector<char>> send(np), recv;
    for (int i = 0; i < n; ++i)
        if (owner[i] == me)
            for (int k = sI[i]; k < sI[i + 1]; ++k)
                if (sJ[k] != me) { Append(send[sJ[k]], key[i]); Append(send[sJ[k]], gid[i]); }
    Exchange(comm, send, recv);
    std::vector<std::unordered_map<int64_t, int64_t>> from(np);
    for (int r = 0; r < np; ++r)
    {
        const int64_t *p = reinterpret_cast<const int64_t *>(recv[r].data());
        const size_t m = recv[r].size() / (2 * sizeof(int64_t));
        from[r].reserve(m);
        for (size_t q = 0; q < m; ++q) from[r].emplace(p[2 * q], p[2 * q + 1]);
    }
    for (int i = 0; i < n; ++i)
        if (owner[i] != me)
        {
            auto it = from[owner[i]].find(key[i]);
            if (it == from[owner[i]].end())
                throw std::runtime_error("pe_par_number_items: rank " + std::to_string(owner[i]) + " did not send the id of shared item with key " +
                                         std::to_string(key[i]) + " (inconsistent sharing tables)");
            gid[i] = it->second;
        }
    if (my_start) *my_start = start;
    if (my_count) *my_count = mine;
    if (global_count) *global_count = total;
    PAR_CATCH
}

extern "C" int pe_par_build_comm_pkg(const pe_host_comm *comm, pe_parcsr_owned *M, const int64_t *col_starts)
{
    PAR_TRY
    const int np = comm->size, me = comm->rank;
    const int64_t my0 = col_starts[me];
    M->recv_procs.clear(); M->recv_vec_starts.assign(1, 0);
    std::vector<std::vector<char>> send(np), recv;
    int cur = -1;
    for (size_t k = 0; k < M->col_map_offd.size(); ++k)
    {
        const int64_t g = M->col_map_offd[k];
        const int o = (int)(std::upper_bound(col_starts, col_starts + np + 1, g) - col_starts) - 1;
        if (o < 0 || o >= np || o == me) throw std::runtime_error("pe_par_build_comm_pkg: ghost column owned by nobody or by the caller");
        if (o != cur)
        {
            if (o < cur) throw std::runtime_error("pe_par_build_comm_pkg: col_map_offd is not ascending");
            if (cur >= 0) M->recv_vec_starts.push_back((int32_t)k);
            M->recv_procs.push_back(o);
            cur = o;
        }
        Append(send[o], g);
    }
    if (cur >= 0) M->recv_vec_starts.push_back((int32_t)M->col_map_offd.size());
    Exchange(comm, send, recv);
    M->send_procs.clear(); M->send_map_starts.assign(1, 0); M->send_map_elmts.clear();
    for (int r = 0; r < np; ++r)
    {
        const size_t m = recv[r].size() / sizeof(int64_t);
        if (!m) continue;
        const int64_t *p = reinterpret_cast<const int64_t *>(recv[r].data());
        M->send_procs.push_back(r);
        for (size_t q = 0; q < m; ++q) M->send_map_elmts.push_back((int32_t)(p[q] - my0));
        M->send_map_starts.push_back((int32_t)M->send_map_elmts.size());
    }
    M->refresh();
    PAR_CATCH
}

namespace
{
struct Trip { int64_t col; double val; };
}

extern "C" int pe_par_assemble(const pe_host_comm *comm, int mode, int32_t nrows, int32_t ncols, const int32_t *I,
                               const int32_t *J, const double *A, const int64_t *row_gid, const int32_t *row_owner,
                               const int64_t *col_gid, const int32_t *col_owner, int64_t row_start, int64_t row_end,
                               int64_t global_rows, int64_t col_start, int64_t col_end, int64_t global_cols,
                               pe_parcsr_owned **out)
{
    (void)ncols; (void)col_owner;
    PAR_TRY
    const int np = comm->size, me = comm->rank;
    const int64_t nown = row_end - row_start;
    // contributions of other ranks to my rows (mode 0)
    std::vector<std::vector<char>> send(np), recv(np);
    if (mode == 0)
    {
        for (int i = 0; i < nrows; ++i)
        {
            const int o = row_owner[i];
            if (o == me) continue;
            const int32_t len = I[i + 1] - I[i];
            if (!len) continue;
            Append(send[o], row_gid[i]);
            const int64_t l64 = len;
            Append(send[o], l64);
            for (int k = I[i]; k < I[i + 1]; ++k) { Append(send[o], col_gid[J[k]]); Append(send[o], A[k]); }
        }
        Exchange(comm, send, recv);
    }
    // count entries per owned true row
    std::vector<int64_t> ptr((size_t)nown + 1, 0);
    for (int i = 0; i < nrows; ++i)
        if (row_owner[i] == me) ptr[(size_t)(row_gid[i] - row_start) + 1] += I[i + 1] - I[i];
    for (int r = 0; r < np; ++r)
    {
        const char *p = recv[r].data(), *e = p + recv[r].size();
        while (p < e)
        {
            int64_t g, len;
            std::memcpy(&g, p, 8); std::memcpy(&len, p + 8, 8);
            if (g < row_start || g >= row_end) throw std::runtime_error("pe_par_assemble: received a row this rank does not own");
            ptr[(size_t)(g - row_start) + 1] += len;
            p += 16 + 16 * len;
        }
    }
    for (int64_t t = 0; t < nown; ++t) ptr[t + 1] += ptr[t];
    std::vector<Trip> ent((size_t)ptr[nown]);
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < nrows; ++i)
        if (row_owner[i] == me)
        {
            int64_t &f = fill[(size_t)(row_gid[i] - row_start)];
            for (int k = I[i]; k < I[i + 1]; ++k) ent[(size_t)f++] = {col_gid[J[k]], A[k]};
        }
    for (int r = 0; r < np; ++r)
    {
        const char *p = recv[r].data(), *e = p + recv[r].size();
        while (p < e)
        {
            int64_t g, len;
            std::memcpy(&g, p, 8); std::memcpy(&len, p + 8, 8);
            p += 16;
            int64_t &f = fill[(size_t)(g - row_start)];
            for (int64_t q = 0; q < len; ++q, p += 16) { Trip t; std::memcpy(&t.col, p, 8); std::memcpy(&t.val, p + 8, 8); ent[(size_t)f++] = t; }
        }
    }
    // sort + combine per row (stable: own contribution first, then ranks ascending), split diag / offd
    auto M = std::make_unique<pe_parcsr_owned>();
    M->diag_i.assign((size_t)nown + 1, 0); M->offd_i.assign((size_t)nown + 1, 0);
    std::vector<int64_t> offd_g;
    for (int64_t t = 0; t < nown; ++t)
    {
        Trip *b = ent.data() + ptr[t], *e = ent.data() + ptr[t + 1];
        std::stable_sort(b, e, [](const Trip &x, const Trip &y) { return x.col < y.col; });
        for (Trip *p = b; p < e;)
        {
            double s = 0.0;
            Trip *q = p;
            while (q < e && q->col == p->col) s += (q++)->val;
            if (p->col >= col_start && p->col < col_end) { M->diag_j.push_back((int32_t)(p->col - col_start)); M->diag_a.push_back(s); }
            else { offd_g.push_back(p->col); M->offd_a.push_back(s); }
            p = q;
        }
        M->diag_i[t + 1] = (int32_t)M->diag_j.size();
        M->offd_i[t + 1] = (int32_t)offd_g.size();
    }
    M->col_map_offd = offd_g;
    std::sort(M->col_map_offd.begin(), M->col_map_offd.end());
    M->col_map_offd.erase(std::unique(M->col_map_offd.begin(), M->col_map_offd.end()), M->col_map_offd.end());
    M->offd_j.resize(offd_g.size());
    for (size_t k = 0; k < offd_g.size(); ++k)
        M->offd_j[k] = (int32_t)(std::lower_bound(M->col_map_offd.begin(), M->col_map_offd.end(), offd_g[k]) - M->col_map_offd.begin());
    M->H.global_num_rows = global_rows; M->H.global_num_cols = global_cols;
    M->H.first_row_index = row_start; M->H.first_col_diag = col_start;
    M->H.num_rows = (int32_t)nown; M->H.num_cols_diag = (int32_t)(col_end - col_start);
    std::vector<int64_t> starts = AllGather<int64_t>(comm, col_start);
    starts.push_back(global_cols);
    M->refresh();
    if (int rc = pe_par_build_comm_pkg(comm, M.get(), starts.data())) return rc;
    *out = M.release();
    PAR_CATCH
}

extern "C" const pe_parcsr_host *pe_parcsr_owned_view(const pe_parcsr_owned *M) { return &M->H; }
extern "C" int pe_parcsr_owned_free(pe_parcsr_owned *M) { delete M; return 0; }
