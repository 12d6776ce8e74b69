// par_host.hpp -- host side of the multi-rank path (one rank <-> one mesh partition <-> one GPU):
//   SharingMap (entity / dof <-> true entity / true dof)   src/structures/SharingMap.cpp:213-1074
//   Assemble / IgnoreNonLocalRange to hypre ParCSR          SharingMap.cpp:930-1011
//   hypre_MatvecCommPkgCreate                               (comm package of a ParCSR matrix)
// All exchanges go through the pe_host_comm callback table (include/parelag_b200_par.h), so
// this file has no MPI / NCCL / CUDA dependency and runs in the CPU-only gloo tests.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "parelag_b200_par.h"

namespace parelag
{
namespace par
{
/// variable-length exchange with every rank (counts first, then payload)
inline void Exchange(const pe_host_comm *comm, const std::vector<std::vector<char>> &send, std::vector<std::vector<char>> &recv)
{
    const int np = comm->size, me = comm->rank;
    std::vector<int64_t> cnt(np), all((size_t)np * np);
    for (int r = 0; r < np; ++r) cnt[r] = (int64_t)send[r].size();
    if (comm->allgather(comm->user, cnt.data(), (int64_t)(sizeof(int64_t) * np), all.data()))
        throw std::runtime_error("host communicator: allgather failed");
    std::vector<int64_t> sb(np), sd(np), rb(np), rd(np);
    int64_t stot = 0, rtot = 0;
    for (int r = 0; r < np; ++r)
    {
        sb[r] = cnt[r]; sd[r] = stot; stot += sb[r];
        rb[r] = all[(size_t)r * np + me]; rd[r] = rtot; rtot += rb[r];
    }
    std::vector<char> sbuf((size_t)std::max<int64_t>(stot, 1)), rbuf((size_t)std::max<int64_t>(rtot, 1));
    for (int r = 0; r < np; ++r) if (sb[r]) std::memcpy(sbuf.data() + sd[r], send[r].data(), (size_t)sb[r]);
    if (comm->alltoallv(comm->user, sbuf.data(), sb.data(), sd.data(), rbuf.data(), rb.data(), rd.data()))
        throw std::runtime_error("host communicator: alltoallv failed");
    recv.assign(np, {});
    for (int r = 0; r < np; ++r) recv[r].assign(rbuf.begin() + rd[r], rbuf.begin() + rd[r] + rb[r]);
}
template <class T> inline void Append(std::vector<char> &b, const T *p, size_t n)
{
    const char *c = reinterpret_cast<const char *>(p);
    b.insert(b.end(), c, c + n * sizeof(T));
}
template <class T> inline void Append(std::vector<char> &b, const T &v) { Append(b, &v, 1); }
template <class T> inline std::vector<T> AllGather(const pe_host_comm *comm, const T &v)
{
    std::vector<T> out(comm->size);
    if (comm->allgather(comm->user, &v, (int64_t)sizeof(T), out.data())) throw std::runtime_error("host communicator: allgather failed");
    return out;
}

/// entity/dof <-> true entity/dof map of one rank (SharingMap): global true id and owner of
/// every local item; owned items are numbered contiguously [start, start + ntrue) in local order
struct SharingMap
{
    std::vector<int64_t> gid, key;
    std::vector<int32_t> owner, sI, sJ;      // sI/sJ: ranks holding each item (CSR, own rank included)
    int64_t start = 0, ntrue = 0, global = 0;
    int rank = 0;
    std::vector<int32_t> true_to_local;   // owned true index -> local index
    int GetLocalSize() const { return (int)gid.size(); }
    int GetTrueLocalSize() const { return (int)ntrue; }
    int64_t GetTrueGlobalSize() const { return global; }
    /// 0: not shared; 1: shared and owned; -1: shared, owned by another rank (SharingMap.cpp:869-890)
    std::vector<int8_t> shared;
    int IsShared(int i) const { return shared.empty() ? 0 : shared[i]; }

    void SetUp(const pe_host_comm *comm, const std::vector<int64_t> &key, const std::vector<int32_t> &sI, const std::vector<int32_t> &sJ)
    {
        const int n = (int)key.size();
        gid.assign(n, -1); owner.assign(n, comm->rank); shared.assign(n, 0);
        this->key = key; this->sI = sI; this->sJ = sJ;
        rank = comm->rank;
        int64_t cnt = 0;
        if (pe_par_number_items(comm, n, key.data(), sI.data(), sJ.data(), gid.data(), owner.data(), &start, &cnt, &global))
            throw std::runtime_error(std::string("SharingMap::SetUp: ") + pe_last_error());
        ntrue = cnt;
        true_to_local.assign((size_t)ntrue, -1);
        for (int i = 0; i < n; ++i)
        {
            if (owner[i] == rank) true_to_local[(size_t)(gid[i] - start)] = i;
            const int ns = sI[i + 1] - sI[i];
            if (ns > 1) shared[i] = owner[i] == rank ? 1 : -1;
        }
    }
    /// SharingMap::Assemble (SharingMap.cpp:768-780): true[t] = sum of the copies of t on all holders
    void Assemble(const pe_host_comm *comm, const double *local_in, double *true_out) const
    {
        const int n = GetLocalSize();
        std::fill(true_out, true_out + ntrue, 0.0);
        std::vector<std::vector<char>> send(comm->size), recv;
        for (int i = 0; i < n; ++i)
        {
            if (owner[i] == rank) true_out[gid[i] - start] += local_in[i];
            else { Append(send[owner[i]], gid[i]); Append(send[owner[i]], local_in[i]); }
        }
        Exchange(comm, send, recv);
        for (int r = 0; r < comm->size; ++r)       // fixed order: ranks ascending
            for (size_t q = 0; q + 16 <= recv[r].size(); q += 16)
            {
                int64_t g; double v;
                std::memcpy(&g, recv[r].data() + q, 8); std::memcpy(&v, recv[r].data() + q + 8, 8);
                true_out[g - start] += v;
            }
    }
    /// SharingMap::Distribute (SharingMap.cpp:664-677): local[i] = true value of item i (owners send to holders)
    void Distribute(const pe_host_comm *comm, const double *true_in, double *local_out) const
    {
        const int n = GetLocalSize();
        std::vector<std::vector<char>> send(comm->size), recv;
        for (int i = 0; i < n; ++i)
            if (owner[i] == rank)
            {
                local_out[i] = true_in[gid[i] - start];
                for (int k = sI[i]; k < sI[i + 1]; ++k)
                    if (sJ[k] != rank) { Append(send[sJ[k]], gid[i]); Append(send[sJ[k]], local_out[i]); }
            }
        Exchange(comm, send, recv);
        std::unordered_map<int64_t, double> got;
        for (int r = 0; r < comm->size; ++r)
            for (size_t q = 0; q + 16 <= recv[r].size(); q += 16)
            {
                int64_t g; double v;
                std::memcpy(&g, recv[r].data() + q, 8); std::memcpy(&v, recv[r].data() + q + 8, 8);
                got[g] = v;
            }
        for (int i = 0; i < n; ++i)
            if (owner[i] != rank)
            {
                auto it = got.find(gid[i]);
                if (it == got.end()) throw std::runtime_error("SharingMap::Distribute: owner did not send a shared item");
                local_out[i] = it->second;
            }
    }
    /// identity map of a single rank
    void SetUpSerial(int n)
    {
        gid.resize(n); owner.assign(n, 0); shared.assign(n, 0); true_to_local.resize(n);
        for (int i = 0; i < n; ++i) { gid[i] = i; true_to_local[i] = i; }
        start = 0; ntrue = global = n; rank = 0;
    }
};
} // namespace par
} // namespace parelag
