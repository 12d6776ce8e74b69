// mfem_compat.hpp -- the minimal slice of MFEM's public interface that ParElag's
// solve path is written against (mfem::Vector / Array / Operator / Solver /
// HypreParMatrix), re-created here because MFEM and hypre are not available in this
// environment.  Names, argument meaning and iterative_mode
// semantics follow MFEM; storage is device-first: a Vector owns an FP64 buffer in HBM
// (pe_vec) with a lazily synchronised host mirror, like MFEM's device memory model
// (HostRead / HostReadWrite / Read / Write).  When real mfem headers are available this
// header is switched off with -DPARELAG_B200_USE_REAL_MFEM.
#pragma once
#ifndef PARELAG_B200_USE_REAL_MFEM
#include <cstdint>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "parelag_b200.h"

namespace parelag
{
/// the process-wide device context (replaces mpi_session, src/utilities/mpiUtils.cpp:23-31)
struct Device
{
    static pe_ctx *&Ctx()
    {
        static pe_ctx *ctx = nullptr;
        return ctx;
    }
    static pe_ctx *Get()
    {
        if (!Ctx())
            throw std::runtime_error("parelag::Device: no device context (create a mpi_session first; "
                                     "there is no CPU fallback)");
        return Ctx();
    }
};

inline void pe_check(int rc, const char *what)
{
    if (rc != 0)
    {
        std::ostringstream os;
        os << what << ": " << pe_last_error();
        throw std::runtime_error(os.str());
    }
}
#define PE_CALL(expr) ::parelag::pe_check((expr), #expr)
} // namespace parelag

namespace mfem
{
template <class T>
class Array
{
public:
    Array() = default;
    explicit Array(int n) : v_(n) {}
    Array(const T *data, int n) : v_(data, data + n) {}
    int Size() const { return (int)v_.size(); }
    void SetSize(int n) { v_.resize(n); }
    void SetSize(int n, const T &val) { v_.assign(n, val); }
    T &operator[](int i) { return v_[i]; }
    const T &operator[](int i) const { return v_[i]; }
    Array &operator=(const T &val)
    {
        for (auto &x : v_) x = val;
        return *this;
    }
    T *GetData() { return v_.data(); }
    const T *GetData() const { return v_.data(); }
    T Max() const
    {
        T m = v_.empty() ? T() : v_[0];
        for (auto &x : v_) if (x > m) m = x;
        return m;
    }
    void Append(const T &x) { v_.push_back(x); }
    typename std::vector<T>::const_iterator begin() const { return v_.begin(); }
    typename std::vector<T>::const_iterator end() const { return v_.end(); }

private:
    std::vector<T> v_;
};

class Vector
{
public:
    Vector() = default;
    explicit Vector(int n) { SetSize(n); }
    Vector(const Vector &o) { *this = o; }
    Vector(Vector &&o) noexcept { swap(o); }
    ~Vector() { Destroy(); }

    Vector &operator=(const Vector &o)
    {
        if (this == &o) return *this;
        SetSize(o.size_);
        if (size_ == 0) return *this;
        if (o.dev_valid_ || !o.host_valid_)
        {
            PE_CALL(pe_vec_copy(o.dev_, dev_));
            dev_valid_ = true; host_valid_ = false;
        }
        else
        {
            host_ = o.host_;
            host_valid_ = true; dev_valid_ = false;
        }
        return *this;
    }
    Vector &operator=(Vector &&o) noexcept
    {
        swap(o);
        return *this;
    }
    /// make this a non-owning alias of another vector's device buffer (the
    /// Vector(data,size) "view" constructor of MFEM, used by Hierarchy::Mult)
    void MakeRef(Vector &base)
    {
        Destroy();
        base.ReadWrite();   // the alias may write: the base's host mirror becomes stale
        dev_ = base.dev_; size_ = base.size_; owns_ = false;
        dev_valid_ = true; host_valid_ = false;
    }
    /// alias of base[offset, offset + n) (mfem::BlockVector::GetBlock)
    void MakeRef(Vector &base, int offset, int n)
    {
        Destroy();
        base.ReadWrite();
        PE_CALL(pe_vec_view(base.dev_, offset, n, &dev_));
        size_ = n; owns_ = true;     // owns the view handle, not the memory
        dev_valid_ = true; host_valid_ = false;
    }
    void MakeRef(const Vector &base, int offset, int n) { MakeRef(const_cast<Vector &>(base), offset, n); }
    void SetSize(int n)
    {
        if (n == size_ && dev_) return;
        Destroy();
        size_ = n;
        PE_CALL(pe_vec_create(parelag::Device::Get(), n, &dev_));
        owns_ = true; dev_valid_ = true; host_valid_ = false;
    }
    int Size() const { return size_; }

    // ---- device access
    const pe_vec *Read() const { SyncDevice(); return dev_; }
    pe_vec *Write() { dev_valid_ = true; host_valid_ = false; return dev_; }
    pe_vec *ReadWrite() { SyncDevice(); host_valid_ = false; return dev_; }
    // ---- host access
    const double *HostRead() const { SyncHost(); return host_.data(); }
    double *HostWrite() { host_.resize(size_); host_valid_ = true; dev_valid_ = false; return host_.data(); }
    double *HostReadWrite() { SyncHost(); dev_valid_ = false; return host_.data(); }
    double *GetData() { return HostReadWrite(); }
    const double *GetData() const { return HostRead(); }
    double &operator()(int i) { return HostReadWrite()[i]; }
    const double &operator()(int i) const { return HostRead()[i]; }

    // ---- arithmetic (device)
    Vector &operator=(double v)
    {
        if (size_) PE_CALL(pe_vec_fill(Write(), v));
        return *this;
    }
    Vector &operator+=(const Vector &o)
    {
        if (size_) PE_CALL(pe_vec_axpby(1.0, o.Read(), 1.0, ReadWrite()));
        return *this;
    }
    Vector &operator-=(const Vector &o)
    {
        if (size_) PE_CALL(pe_vec_axpby(-1.0, o.Read(), 1.0, ReadWrite()));
        return *this;
    }
    Vector &operator*=(double a)
    {
        if (size_) PE_CALL(pe_vec_scale(ReadWrite(), a));
        return *this;
    }
    /// *this = a * x
    Vector &Set(double a, const Vector &x)
    {
        SetSize(x.Size());
        if (size_) PE_CALL(pe_vec_axpby(a, x.Read(), 0.0, Write()));
        return *this;
    }
    /// *this += a * x
    Vector &Add(double a, const Vector &x)
    {
        if (size_) PE_CALL(pe_vec_axpby(a, x.Read(), 1.0, ReadWrite()));
        return *this;
    }
    double operator*(const Vector &o) const
    {
        double out = 0.0;
        PE_CALL(pe_vec_dot(Read(), o.Read(), &out));
        return out;
    }
    void swap(Vector &o) noexcept
    {
        std::swap(dev_, o.dev_); std::swap(size_, o.size_); std::swap(owns_, o.owns_);
        std::swap(host_, o.host_); std::swap(host_valid_, o.host_valid_); std::swap(dev_valid_, o.dev_valid_);
    }

private:
    void Destroy()
    {
        if (dev_ && owns_) pe_vec_free(dev_);
        dev_ = nullptr; size_ = 0; owns_ = true; host_valid_ = false; dev_valid_ = false;
        host_.clear();
    }
    void SyncDevice() const
    {
        if (!dev_valid_ && host_valid_ && size_) PE_CALL(pe_vec_upload(dev_, host_.data()));
        dev_valid_ = true;
    }
    void SyncHost() const
    {
        if (!host_valid_)
        {
            host_.resize(size_);
            if (size_ && dev_valid_) PE_CALL(pe_vec_download(dev_, host_.data()));
            host_valid_ = true;
        }
    }
    pe_vec *dev_ = nullptr;
    int size_ = 0;
    bool owns_ = true;
    mutable std::vector<double> host_;
    mutable bool host_valid_ = false, dev_valid_ = false;
};

/// z = a*x + b*y  (mfem::add)
inline void add(double a, const Vector &x, double b, const Vector &y, Vector &z)
{
    if (z.Size()) PE_CALL(pe_vec_add3(a, x.Read(), b, y.Read(), z.Write()));
}
/// z = x + a*y  (mfem::add)
inline void add(const Vector &x, double a, const Vector &y, Vector &z)
{
    if (z.Size()) PE_CALL(pe_vec_add3(1.0, x.Read(), a, y.Read(), z.Write()));
}

class Operator
{
public:
    explicit Operator(int s = 0) : height(s), width(s) {}
    Operator(int h, int w) : height(h), width(w) {}
    virtual ~Operator() = default;
    int Height() const { return height; }
    int Width() const { return width; }
    int NumRows() const { return height; }
    int NumCols() const { return width; }
    virtual void Mult(const Vector &x, Vector &y) const = 0;
    virtual void MultTranspose(const Vector &, Vector &) const
    {
        throw std::logic_error("Operator::MultTranspose() is not overloaded!");
    }

protected:
    int height, width;
};

class Solver : public Operator
{
public:
    explicit Solver(int s = 0, bool iter_mode = false) : Operator(s), iterative_mode(iter_mode) {}
    Solver(int h, int w, bool iter_mode = false) : Operator(h, w), iterative_mode(iter_mode) {}
    /// If true, use the second argument of Mult() as an initial guess.
    bool iterative_mode;
    virtual void SetOperator(const Operator &op) = 0;
};

/// Device-resident ParCSR matrix behind MFEM's HypreParMatrix interface.
class HypreParMatrix : public Operator
{
public:
    /// takes ownership of a device matrix
    explicit HypreParMatrix(pe_mat *A) : A_(A)
    {
        int32_t nr, nc;
        PE_CALL(pe_mat_info(A_, &nr, &nc, nullptr, &nnz_diag_, &nnz_offd_));
        height = nr; width = nc;
    }
    /// upload a host ParCSR description (what hypreExtension hands over)
    explicit HypreParMatrix(const pe_parcsr_host &H)
    {
        PE_CALL(pe_mat_upload(parelag::Device::Get(), &H, &A_));
        int32_t nr, nc;
        PE_CALL(pe_mat_info(A_, &nr, &nc, nullptr, &nnz_diag_, &nnz_offd_));
        height = nr; width = nc;
    }
    ~HypreParMatrix() override { pe_mat_free(A_); }
    HypreParMatrix(const HypreParMatrix &) = delete;
    HypreParMatrix &operator=(const HypreParMatrix &) = delete;

    void Mult(const Vector &x, Vector &y) const override
    {
        PE_CALL(pe_spmv(parelag::Device::Get(), 1.0, A_, x.Read(), 0.0, y.Write()));
    }
    void Mult(double a, const Vector &x, double b, Vector &y) const
    {
        PE_CALL(pe_spmv(parelag::Device::Get(), a, A_, x.Read(), b, b == 0.0 ? y.Write() : y.ReadWrite()));
    }
    void MultTranspose(const Vector &x, Vector &y) const override
    {
        PE_CALL(pe_spmv_t(parelag::Device::Get(), 1.0, A_, x.Read(), 0.0, y.Write()));
    }
    int64_t M() const { return height; }
    int64_t N() const { return width; }
    int64_t NNZ() const { return nnz_diag_ + nnz_offd_; }
    pe_mat *Handle() const { return A_; }
    /// give up ownership of the device matrix
    pe_mat *Release() { pe_mat *a = A_; A_ = nullptr; return a; }

private:
    pe_mat *A_ = nullptr;
    int64_t nnz_diag_ = 0, nnz_offd_ = 0;
};

/// mfem::RAP(A,P) -> hypre_BoomerAMGBuildCoarseOperator (Hierarchy.cpp:365)
inline HypreParMatrix *RAP(const HypreParMatrix *A, const HypreParMatrix *P)
{
    pe_mat *Ac = nullptr;
    PE_CALL(pe_rap(parelag::Device::Get(), nullptr, A->Handle(), P->Handle(), &Ac));
    return new HypreParMatrix(Ac);
}
inline HypreParMatrix *RAP(const HypreParMatrix *Rt, const HypreParMatrix *A, const HypreParMatrix *P)
{
    pe_mat *Ac = nullptr;
    PE_CALL(pe_rap(parelag::Device::Get(), Rt->Handle(), A->Handle(), P->Handle(), &Ac));
    return new HypreParMatrix(Ac);
}
} // namespace mfem

/// hypre_ParCSRMatrixFixZeroRows on the device matrix (Hierarchy.cpp:366-371)
inline int hypre_ParCSRMatrixFixZeroRows(mfem::HypreParMatrix &A)
{
    int32_t nfixed = 0;
    PE_CALL(pe_fix_zero_rows(parelag::Device::Get(), A.Handle(), &nfixed));
    return nfixed;
}
#endif // PARELAG_B200_USE_REAL_MFEM
