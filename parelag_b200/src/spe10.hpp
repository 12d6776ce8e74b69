// spe10.hpp -- the SPE10 permeability data set as the reference reads and evaluates it
// (src/SPE10/InversePermeabilityFunction.{hpp,cpp}; examples/MultigridTestSPE10.cpp:85-187,377-395):
//   * file layout: three blocks K_x, K_y, K_z of 60 x 220 x 85 numbers each, x fastest, then y, then z
//     (InversePermeabilityFunction.cpp:57-100); a sub-box Nx x Ny x Nz of the data set is cut out while reading;
//     what is stored is 1 / K, component after component;
//   * evaluation at a point (:141-178): cell i = Nx - 1 - floor(x / hx / (1 + 3e-16)), j = floor(y / hy / (1 + 3e-16)),
//     k = Nz - 1 - floor(z / hz / (1 + 3e-16)) -- the x and z axes of the data set run against the mesh axes;
//     2-d slices fix one index.
// The coefficient enters the H(div) element mass matrices as the diagonal tensor diag(1 / K_x, 1 / K_y, 1 / K_z)
// (VectorFunctionCoefficient in VectorFEMassIntegrator); on axis-aligned hexahedra that scales the three 2 x 2 axis blocks
// of the RT0 mass matrix (amge_hex.hpp, beta with three components per element).
// Same class and method names as the reference; mfem::Vector arguments are plain pointers (host-only code).
#pragma once
#include <algorithm>
#include <cmath>
#include <fstream>
#include <string>
#include <vector>
#include "parelag_b200_par.h"
#include "parelag_core.hpp"

namespace parelag
{
class InversePermeabilityFunction
{
public:
    enum SliceOrientation { NONE, XY, XZ, YZ };
    static void SetNumberCells(int nx, int ny, int nz) { st().Nx = nx; st().Ny = ny; st().Nz = nz; }
    static void SetMeshSizes(double hx, double hy, double hz) { st().hx = hx; st().hy = hy; st().hz = hz; }
    static void Set2DSlice(SliceOrientation o, int npos) { st().orientation = o; st().npos = npos; }
    static void SetConstantInversePermeability(double ipx, double ipy, double ipz)
    {
        State &s = st();
        const size_t n = (size_t)s.Nx * s.Ny * s.Nz;
        s.ip.resize(3 * n);
        std::fill(s.ip.begin(), s.ip.begin() + n, ipx);
        std::fill(s.ip.begin() + n, s.ip.begin() + 2 * n, ipy);
        std::fill(s.ip.begin() + 2 * n, s.ip.end(), ipz);
    }
    /// reads the Nx x Ny x Nz corner of the 60 x 220 x 85 data set, all three components, and stores the reciprocals
    static void ReadPermeabilityFile(const std::string &fileName)
    {
        State &s = st();
        std::ifstream in(fileName);
        PARELAG_TEST_FOR_EXCEPTION(!in, std::runtime_error, "InversePermeabilityFunction::ReadPermeabilityFile: cannot open " << fileName);
        PARELAG_TEST_FOR_EXCEPTION(s.Nx < 1 || s.Nx > FX || s.Ny < 1 || s.Ny > FY || s.Nz < 1 || s.Nz > FZ, std::runtime_error,
                                   "InversePermeabilityFunction: the data set has " << FX << " x " << FY << " x " << FZ << " cells");
        const size_t n = (size_t)s.Nx * s.Ny * s.Nz;
        s.ip.assign(3 * n, 0.0);
        double v;
        auto skip = [&](long count) { for (long q = 0; q < count; ++q) in >> v; };
        size_t pos = 0;
        for (int comp = 0; comp < 3; ++comp)
        {
            for (int k = 0; k < s.Nz; ++k)
            {
                for (int j = 0; j < s.Ny; ++j)
                {
                    for (int i = 0; i < s.Nx; ++i) { in >> v; s.ip[pos++] = 1.0 / v; }
                    skip(FX - s.Nx);                                   // rest of the row
                }
                skip((long)(FY - s.Ny) * FX);                          // rest of the layer
            }
            if (comp < 2) skip((long)(FZ - s.Nz) * FY * FX);           // rest of the component
        }
        PARELAG_TEST_FOR_EXCEPTION(!in, std::runtime_error, "InversePermeabilityFunction::ReadPermeabilityFile: " << fileName << " ends early");
    }
    /// the reference's ReadPermeabilityFile(fileName, comm): rank 0 reads, everybody receives (MPI_Bcast there; the
    /// host communicator's allgather here)
    static void ReadPermeabilityFile(const std::string &fileName, const pe_host_comm *comm)
    {
        State &s = st();
        if (!comm || comm->size <= 1) { ReadPermeabilityFile(fileName); return; }
        const size_t n = (size_t)3 * s.Nx * s.Ny * s.Nz;
        if (comm->rank == 0) ReadPermeabilityFile(fileName); else s.ip.assign(n, 0.0);
        std::vector<double> all(n * (size_t)comm->size);
        PARELAG_TEST_FOR_EXCEPTION(comm->allgather(comm->user, s.ip.data(), (int64_t)(n * sizeof(double)), all.data()) != 0, std::runtime_error,
                                   "InversePermeabilityFunction::ReadPermeabilityFile: host allgather failed");
        std::copy(all.begin(), all.begin() + n, s.ip.begin());
    }
    template <class F> static void Transform(const F &f) { for (double &v : st().ip) v = f(v); }
    /// val[0..2] (val[0..1] on a slice) = 1 / K at the point x
    static void InversePermeability(const double *x, double *val)
    {
        const State &s = st();
        int i = 0, j = 0, k = 0;
        Cell(x, i, j, k);
        const size_t n = (size_t)s.Nx * s.Ny * s.Nz, c = (size_t)s.Ny * s.Nx * k + (size_t)s.Nx * j + i;
        PARELAG_TEST_FOR_EXCEPTION(s.ip.size() != 3 * n || c >= n, std::runtime_error, "InversePermeabilityFunction: no data for this point");
        val[0] = s.ip[c];
        val[1] = s.ip[c + n];
        if (s.orientation == NONE) val[2] = s.ip[c + 2 * n];
    }
    static void NegativeInversePermeability(const double *x, double *val) { InversePermeability(x, val); for (int q = 0; q < Dim(); ++q) val[q] = -val[q]; }
    static void Permeability(const double *x, double *val) { InversePermeability(x, val); for (int q = 0; q < Dim(); ++q) val[q] = 1.0 / val[q]; }
    /// diagonal tensor K (row major dim x dim)
    static void PermeabilityTensor(const double *x, double *val)
    {
        double k[3];
        Permeability(x, k);
        const int d = Dim();
        std::fill(val, val + d * d, 0.0);
        for (int q = 0; q < d; ++q) val[q * d + q] = k[q];
    }
    static double PermeabilityXY(const double *x)
    {
        const State &s = st();
        const int i = s.Nx - 1 - (int)std::floor(x[0] / s.hx / (1. + 3e-16)), j = (int)std::floor(x[1] / s.hy / (1. + 3e-16));
        return 1.0 / s.ip[(size_t)s.Ny * s.Nx * s.npos + (size_t)s.Nx * j + i];
    }
    static double Norm2InversePermeability(const double *x) { double v[3]; InversePermeability3(x, v); return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    static double Norm1InversePermeability(const double *x) { double v[3]; InversePermeability3(x, v); return std::fabs(v[0]) + std::fabs(v[1]) + std::fabs(v[2]); }
    static double NormInfInversePermeability(const double *x) { double v[3]; InversePermeability3(x, v); return std::max(std::fabs(v[0]), std::max(std::fabs(v[1]), std::fabs(v[2]))); }
    static double InvNorm2(const double *x) { return 1.0 / Norm2InversePermeability(x); }
    static double InvNorm1(const double *x) { return 1.0 / Norm1InversePermeability(x); }
    static double InvNormInf(const double *x) { return 1.0 / NormInfInversePermeability(x); }
    static void ClearMemory() { std::vector<double>().swap(st().ip); }
    static const std::vector<double> &Data() { return st().ip; }
    static bool Loaded() { const State &s = st(); return s.ip.size() == (size_t)3 * s.Nx * s.Ny * s.Nz && !s.ip.empty(); }

private:
    static constexpr int FX = 60, FY = 220, FZ = 85;      // cells of the full data set
    struct State
    {
        int Nx = FX, Ny = FY, Nz = FZ;
        double hx = 20, hy = 10, hz = 2;
        std::vector<double> ip;
        SliceOrientation orientation = NONE;
        int npos = -1;
    };
    static State &st() { static State s; return s; }
    static int Dim() { return st().orientation == NONE ? 3 : 2; }
    static void InversePermeability3(const double *x, double *v) { v[2] = 0.0; InversePermeability(x, v); }
    static void Cell(const double *x, int &i, int &j, int &k)
    {
        const State &s = st();
        const double g = 1. + 3e-16;
        switch (s.orientation)
        {
        case NONE: i = s.Nx - 1 - (int)std::floor(x[0] / s.hx / g); j = (int)std::floor(x[1] / s.hy / g); k = s.Nz - 1 - (int)std::floor(x[2] / s.hz / g); break;
        case XY: i = s.Nx - 1 - (int)std::floor(x[0] / s.hx / g); j = (int)std::floor(x[1] / s.hy / g); k = s.npos; break;
        case XZ: i = s.Nx - 1 - (int)std::floor(x[0] / s.hx / g); j = s.npos; k = s.Nz - 1 - (int)std::floor(x[2] / s.hz / g); break;
        case YZ: i = s.npos; j = (int)std::floor(x[1] / s.hy / g); k = s.Nz - 1 - (int)std::floor(x[2] / s.hz / g); break;
        }
    }
};
} // namespace parelag
