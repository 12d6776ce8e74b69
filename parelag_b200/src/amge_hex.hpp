// amge_hex.hpp -- synthetic input producer: the fine-level de Rham sequence
// (DeRhamSequence3D_FE, src/amge/DeRhamSequenceFE.cpp:633-722, at lowest order) on an
// axis-aligned structured hexahedral mesh, with MFEM's dof conventions (H1: vertex
// values, Nedelec: edge circulations, Raviart-Thomas: face fluxes, L2: cell values),
// exact element / facet / ridge mass matrices, PV-trace geometry, and the order-0
// upscaling targets of SetUpscalingTargets (DeRhamSequenceFE.cpp:927-982).
// Numbering (x fastest): element (i,j,k) -> i + nx*(j + ny*k); facets: x-normal faces,
// then y-, then z-normal; ridges: x-, y-, z-edges; peaks: vertices.  Every facet / ridge
// is oriented along its positive axis.  Boundary attributes as mfem::Mesh::Make3D
// (z=0:1, y=0:2, x=L:3, y=L:4, x=0:5, z=L:6).  Same conventions as oracle/amge.py.
// Optional vertex coordinates turn the cells into trilinear hexahedra (examples/3DHdivWeakScaling.cpp:148-158);
// the H(div)-L2 part of the sequence (forms 2 and 3) is then built by quadrature, see BuildFineHexSequenceDeformed.
#pragma once
#include "amge_dofs.hpp"

namespace parelag
{
struct StructuredHexMesh
{
    int nx, ny, nz;
    double hx, hy, hz;
    /// multi-rank: this mesh is one box of a larger one.  iface[2*axis+side]: that box face is an
    /// interface with another rank -- its facets carry the pseudo boundary attribute 7 + 2*axis + side
    /// (never essential; it only makes the minimal intersection sets split there, amge_par.hpp);
    /// x0/y0/z0: origin of the box
    bool iface[6] = {false, false, false, false, false, false};
    double x0 = 0.0, y0 = 0.0, z0 = 0.0;
    /// optional: moved vertices (nv x 3, row-major, index-grid numbering) -> trilinear hexahedra
    std::vector<double> coords;
    bool deformed() const { return !coords.empty(); }
    StructuredHexMesh(int nx_, int ny_, int nz_, double Lx = 1.0, double Ly = 1.0, double Lz = 1.0)
        : nx(nx_), ny(ny_), nz(nz_), hx(Lx / nx_), hy(Ly / ny_), hz(Lz / nz_) {}
    bool any_interface() const { for (bool b : iface) if (b) return true; return false; }
    int64_t nel() const { return (int64_t)nx * ny * nz; }
    int nfx() const { return (nx + 1) * ny * nz; }
    int nfy() const { return nx * (ny + 1) * nz; }
    int nfz() const { return nx * ny * (nz + 1); }
    int nf() const { return nfx() + nfy() + nfz(); }
    int nex() const { return nx * (ny + 1) * (nz + 1); }
    int ney() const { return (nx + 1) * ny * (nz + 1); }
    int nez() const { return (nx + 1) * (ny + 1) * nz; }
    int ne() const { return nex() + ney() + nez(); }
    int nv() const { return (nx + 1) * (ny + 1) * (nz + 1); }
    int el(int i, int j, int k) const { return i + nx * (j + ny * k); }
    int fx(int i, int j, int k) const { return i + (nx + 1) * (j + ny * k); }
    int fy(int i, int j, int k) const { return nfx() + i + nx * (j + (ny + 1) * k); }
    int fz(int i, int j, int k) const { return nfx() + nfy() + i + nx * (j + ny * k); }
    int ex(int i, int j, int k) const { return i + nx * (j + (ny + 1) * k); }
    int ey(int i, int j, int k) const { return nex() + i + (nx + 1) * (j + ny * k); }
    int ez(int i, int j, int k) const { return nex() + ney() + i + (nx + 1) * (j + (ny + 1) * k); }
    int vx(int i, int j, int k) const { return i + (nx + 1) * (j + (ny + 1) * k); }

    static void push_row(HostCSR &M, std::vector<std::pair<int, double>> &ent)
    {
        std::sort(ent.begin(), ent.end());
        for (auto &p : ent) { M.J.push_back(p.first); M.A.push_back(p.second); }
        M.I.push_back((int)M.J.size());
        ent.clear();
    }
    std::shared_ptr<AgglomeratedTopology> Topology() const
    {
        std::vector<std::pair<int, double>> r;
        HostCSR B0, B1, B2, fb;
        B0.nrows = (int)nel(); B0.ncols = nf(); B0.I = {0};
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)
        {
            r = {{fx(i, j, k), -1.0}, {fx(i + 1, j, k), 1.0}, {fy(i, j, k), -1.0}, {fy(i, j + 1, k), 1.0}, {fz(i, j, k), -1.0}, {fz(i, j, k + 1), 1.0}};
            push_row(B0, r);
        }
        B1.nrows = nf(); B1.ncols = ne(); B1.I = {0};
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i <= nx; ++i)   // x-faces: dEz/dy - dEy/dz
        { r = {{ez(i, j + 1, k), 1.0}, {ez(i, j, k), -1.0}, {ey(i, j, k + 1), -1.0}, {ey(i, j, k), 1.0}}; push_row(B1, r); }
        for (int k = 0; k < nz; ++k) for (int j = 0; j <= ny; ++j) for (int i = 0; i < nx; ++i)   // y-faces: dEx/dz - dEz/dx
        { r = {{ex(i, j, k + 1), 1.0}, {ex(i, j, k), -1.0}, {ez(i + 1, j, k), -1.0}, {ez(i, j, k), 1.0}}; push_row(B1, r); }
        for (int k = 0; k <= nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)   // z-faces: dEy/dx - dEx/dy
        { r = {{ey(i + 1, j, k), 1.0}, {ey(i, j, k), -1.0}, {ex(i, j + 1, k), -1.0}, {ex(i, j, k), 1.0}}; push_row(B1, r); }
        B2.nrows = ne(); B2.ncols = nv(); B2.I = {0};
        for (int k = 0; k <= nz; ++k) for (int j = 0; j <= ny; ++j) for (int i = 0; i < nx; ++i)
        { r = {{vx(i, j, k), -1.0}, {vx(i + 1, j, k), 1.0}}; push_row(B2, r); }
        for (int k = 0; k <= nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i <= nx; ++i)
        { r = {{vx(i, j, k), -1.0}, {vx(i, j + 1, k), 1.0}}; push_row(B2, r); }
        for (int k = 0; k < nz; ++k) for (int j = 0; j <= ny; ++j) for (int i = 0; i <= nx; ++i)
        { r = {{vx(i, j, k), -1.0}, {vx(i, j, k + 1), 1.0}}; push_row(B2, r); }
        fb.nrows = nf(); fb.ncols = any_interface() ? 12 : 6;
        std::vector<int> attr(nf(), -1);
        const int ax0 = iface[0] ? 6 : 4, ax1 = iface[1] ? 7 : 2, ay0 = iface[2] ? 8 : 1, ay1 = iface[3] ? 9 : 3;
        const int az0 = iface[4] ? 10 : 0, az1 = iface[5] ? 11 : 5;
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) { attr[fx(0, j, k)] = ax0; attr[fx(nx, j, k)] = ax1; }
        for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) { attr[fy(i, 0, k)] = ay0; attr[fy(i, ny, k)] = ay1; }
        for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) { attr[fz(i, j, 0)] = az0; attr[fz(i, j, nz)] = az1; }
        fb.I = {0};
        for (int f = 0; f < nf(); ++f)
        {
            if (attr[f] >= 0) { fb.J.push_back(attr[f]); fb.A.push_back(1.0); }
            fb.I.push_back((int)fb.J.size());
        }
        std::vector<HostCSR> B;
        B.push_back(std::move(B0)); B.push_back(std::move(B1)); B.push_back(std::move(B2));
        return std::make_shared<AgglomeratedTopology>(std::move(B), std::move(fb), 3);
    }
};

namespace hexfe
{
inline void kron2(const double *A, int na, const double *B, int nb, double *C)   // C = A (x) B, row-major
{
    const int n = na * nb;
    for (int ia = 0; ia < na; ++ia) for (int ib = 0; ib < nb; ++ib)
        for (int ja = 0; ja < na; ++ja) for (int jb = 0; jb < nb; ++jb)
            C[(ia * nb + ib) * n + (ja * nb + jb)] = A[ia * na + ja] * B[ib * nb + jb];
}
/// place the s-scaled m x m block B at (o,o) of the n x n row-major matrix C
inline void place(double *C, int n, int o, const double *B, int m, double s)
{
    for (int a = 0; a < m; ++a) for (int b = 0; b < m; ++b) C[(o + a) * n + (o + b)] = s * B[a * m + b];
}
inline void fill_pool(BlockPool &P, int count, const double *blk, int m, const double *w = nullptr)
{
    // one allocation per pool (the fine element pools are gigabytes at 144^3 hexahedra), filled in parallel
    const size_t n0 = P.size.size(), v0 = P.vals.size(), mm = (size_t)m * m;
    P.size.resize(n0 + (size_t)count, m);
    P.off.resize(n0 + (size_t)count + 1);
    P.vals.resize(v0 + (size_t)count * mm);
    double *vals = P.vals.data() + v0;
    int64_t *off = P.off.data() + n0 + 1;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < count; ++e)
    {
        off[e] = (int64_t)(v0 + (size_t)(e + 1) * mm);
        const double s = w ? w[e] : 1.0;
        double *d = vals + (size_t)e * mm;
        for (size_t q = 0; q < mm; ++q) d[q] = s * blk[q];
    }
}
} // namespace hexfe

/// DeRhamSequence3D_FE at lowest order on trilinear hexahedra (all four forms)
/// (DeRhamSequenceFE.cpp:633-684).  Per cell: volume (MassIntegrator on P0: 2-point Gauss rule, exact), RT0 mass
/// matrix (VectorFEMassIntegrator, contravariant Piola map v = J vhat / det J, Gauss rule of order OrderW + 2 = 4:
/// 3 points per direction), D_2 = net flux / volume (DivergenceInterpolator2, bilinIntegrators.hpp:272-290).
/// Per facet: N = dr/du x dr/dv at the centre, oriented along the +index axis; trace mass 1/|N|
/// (VolumetricFEMassIntegrator, one-point rule), PV-trace weight |N| (InterpolatePV_HdivTraces,
/// DeRhamSequenceFE.cpp:810-857), targets = fluxes N.e_c of the constant fields (RT_HexahedronElement::Project).
/// Local dof order of a cell: x-, x+, y-, y+, z-, z+ (ascending facet id), every basis function with unit flux
/// along the +index axis.  With jformStart = 1 also the Nedelec part: element mass matrices (covariant Piola map
/// w = J^-T what, same Gauss rule; local order = ascending edge id: x-edges (j,k),(j+1,k),(j,k+1),(j+1,k+1), y-edges
/// (i,k),(i+1,k),(i,k+1),(i+1,k+1), z-edges (i,j),(i+1,j),(i,j+1),(i+1,j+1)), facet mass matrices of the tangential
/// traces (ND_3D_FacetMassIntegrator, bilinIntegrators.cpp:106-157, 2x2 Gauss rule; local order: the two edges along the
/// first in-plane axis, then the two along the second; in-plane axes x-faces (y,z), y-faces (x,z), z-faces (x,y)), edge
/// masses 1/|t|, PV-trace weights |t| and circulation targets t.e_c with t = end point - start point.
/// With jformStart = 0 also the H1 part: trilinear nodal mass matrices of the cells (3-point rule) and facets (2-point
/// rule), exact edge mass matrices |t| [1/3 1/6; 1/6 1/3], unit vertex masses, targets 1, z, y, x at the moved vertices.
/// Same arithmetic as oracle/amge.py:DeformedHexMesh.
inline void BuildFineHexSequenceDeformed(const StructuredHexMesh &mesh, const std::shared_ptr<AgglomeratedTopology> &topo,
                                         const double *alpha, const double *beta, int jstart, SequenceData &S, std::vector<HostCSR> &D)
{
    PARELAG_TEST_FOR_EXCEPTION((int64_t)mesh.coords.size() != (int64_t)3 * mesh.nv(), std::runtime_error,
                               "vertex coordinates: expected nv x 3 values");
    using hexfe::fill_pool;
    const int nel = (int)mesh.nel(), nf = mesh.nf();
    S.topo = topo; S.nforms = 4; S.jstart = jstart; S.is_fe = true;
    S.dof.resize(4);
    for (int j = 0; j < 4; ++j)
    {
        auto dh = std::make_shared<DofHandlerX>(3 - j, topo);
        dh->ndofs = topo->GetNumberLocalEntities(3 - j);
        for (int c = 0; c <= 3 - j; ++c)
            dh->entity_dof[c] = (c == 3 - j) ? hostcsr::Identity(topo->GetNumberLocalEntities(c)) : topo->GetConnectivity(c, 3 - j);
        dh->ComputeBoundaryMask();
        S.dof[j] = dh;
    }
    const double *X = mesh.coords.data();
    auto vtx = [&](int i, int j, int k) { return X + (size_t)3 * mesh.vx(i, j, k); };
    const double s35 = std::sqrt(0.6), s13 = 1.0 / std::sqrt(3.0);
    const double g3[3] = {0.5 * (1.0 - s35), 0.5, 0.5 * (1.0 + s35)}, w3[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
    const double g2[2] = {0.5 * (1.0 - s13), 0.5 * (1.0 + s13)}, w2[2] = {0.5, 0.5};
    std::vector<double> vol(nel, 0.0);
    BlockPool &M30 = S.M[{3, 0}], &M20 = S.M[{2, 0}], &M21 = S.M[{2, 1}];
    {
        double one = 1.0;
        fill_pool(M30, nel, &one, 1);
        double zero36[36] = {0};
        fill_pool(M20, nel, zero36, 6);
        fill_pool(M21, nf, &one, 1);
    }
    const bool with_curl = jstart <= 1, with_h1 = jstart <= 0;
    double *m00 = nullptr;
    if (with_h1)
    {
        double z64[64] = {0};
        fill_pool(S.M[{0, 0}], nel, z64, 8);
        m00 = S.M[{0, 0}].vals.data();
    }
    double *m10 = nullptr;
    if (with_curl)
    {
        std::vector<double> z144(144, 0.0);
        fill_pool(S.M[{1, 0}], nel, z144.data(), 12);
        m10 = S.M[{1, 0}].vals.data();
    }
    double *m30 = M30.vals.data(), *m20 = M20.vals.data(), *m21 = M21.vals.data();
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nel; ++e)
    {
        const int i = e % mesh.nx, j = (e / mesh.nx) % mesh.ny, k = e / (mesh.nx * mesh.ny);
        const double *c[2][2][2];
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int cc = 0; cc < 2; ++cc) c[a][b][cc] = vtx(i + a, j + b, k + cc);
        auto jac = [&](double xh, double yh, double zh, double *rx, double *ry, double *rz) {
            const double sx[2] = {1 - xh, xh}, sy[2] = {1 - yh, yh}, sz[2] = {1 - zh, zh}, d[2] = {-1.0, 1.0};
            for (int t = 0; t < 3; ++t) rx[t] = ry[t] = rz[t] = 0.0;
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int cc = 0; cc < 2; ++cc)
                for (int t = 0; t < 3; ++t)
                {
                    const double x = c[a][b][cc][t];
                    rx[t] += x * (d[a] * sy[b] * sz[cc]);
                    ry[t] += x * (sx[a] * d[b] * sz[cc]);
                    rz[t] += x * (sx[a] * sy[b] * d[cc]);
                }
        };
        auto det3 = [](const double *rx, const double *ry, const double *rz) {
            return rx[0] * (ry[1] * rz[2] - ry[2] * rz[1]) + rx[1] * (ry[2] * rz[0] - ry[0] * rz[2]) + rx[2] * (ry[0] * rz[1] - ry[1] * rz[0]);
        };
        double rx[3], ry[3], rz[3], v = 0.0;
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int cc = 0; cc < 2; ++cc)
        {
            jac(g2[a], g2[b], g2[cc], rx, ry, rz);
            v += w2[a] * w2[b] * w2[cc] * det3(rx, ry, rz);
        }
        vol[e] = v;
        m30[e] = v * (alpha ? alpha[e] : 1.0);
        double Me[36] = {0};
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int cc = 0; cc < 3; ++cc)
        {
            jac(g3[a], g3[b], g3[cc], rx, ry, rz);
            const double det = det3(rx, ry, rz), w = w3[a] * w3[b] * w3[cc];
            double V[6][3];
            for (int t = 0; t < 3; ++t)
            {
                V[0][t] = rx[t] * (1 - g3[a]); V[1][t] = rx[t] * g3[a];
                V[2][t] = ry[t] * (1 - g3[b]); V[3][t] = ry[t] * g3[b];
                V[4][t] = rz[t] * (1 - g3[cc]); V[5][t] = rz[t] * g3[cc];
            }
            for (int p = 0; p < 6; ++p) for (int q = 0; q < 6; ++q)
                Me[p * 6 + q] += w * (V[p][0] * V[q][0] + V[p][1] * V[q][1] + V[p][2] * V[q][2]) / det;
            if (with_curl)
            {
                // rows of J^-1 = gradients of the reference coordinates: J^-T what = what_x grad(xh) + ...
                const double gx[3] = {(ry[1] * rz[2] - ry[2] * rz[1]) / det, (ry[2] * rz[0] - ry[0] * rz[2]) / det, (ry[0] * rz[1] - ry[1] * rz[0]) / det};
                const double gy[3] = {(rz[1] * rx[2] - rz[2] * rx[1]) / det, (rz[2] * rx[0] - rz[0] * rx[2]) / det, (rz[0] * rx[1] - rz[1] * rx[0]) / det};
                const double gz[3] = {(rx[1] * ry[2] - rx[2] * ry[1]) / det, (rx[2] * ry[0] - rx[0] * ry[2]) / det, (rx[0] * ry[1] - rx[1] * ry[0]) / det};
                const double pa[2] = {1 - g3[a], g3[a]}, pb[2] = {1 - g3[b], g3[b]}, pc[2] = {1 - g3[cc], g3[cc]};
                const int pr[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
                double Wv[12][3];
                for (int t4 = 0; t4 < 4; ++t4)
                    for (int t = 0; t < 3; ++t)
                    {
                        Wv[t4][t] = pb[pr[t4][0]] * pc[pr[t4][1]] * gx[t];
                        Wv[4 + t4][t] = pa[pr[t4][0]] * pc[pr[t4][1]] * gy[t];
                        Wv[8 + t4][t] = pa[pr[t4][0]] * pb[pr[t4][1]] * gz[t];
                    }
                double *Mn = m10 + (size_t)e * 144;
                for (int p = 0; p < 12; ++p) for (int q = 0; q < 12; ++q)
                    Mn[p * 12 + q] += w * det * (Wv[p][0] * Wv[q][0] + Wv[p][1] * Wv[q][1] + Wv[p][2] * Wv[q][2]);
            }
            if (with_h1)
            {
                const double pa[2] = {1 - g3[a], g3[a]}, pb[2] = {1 - g3[b], g3[b]}, pc[2] = {1 - g3[cc], g3[cc]};
                double sh[8];
                for (int r = 0; r < 2; ++r) for (int q = 0; q < 2; ++q) for (int p = 0; p < 2; ++p) sh[4 * r + 2 * q + p] = pa[p] * pb[q] * pc[r];
                double *Mh = m00 + (size_t)e * 64;
                for (int p = 0; p < 8; ++p) for (int q = 0; q < 8; ++q) Mh[p * 8 + q] += w * det * sh[p] * sh[q];
            }
        }
        const double be = beta ? beta[e] : 1.0;
        for (int q = 0; q < 36; ++q) m20[(size_t)e * 36 + q] = be * Me[q];
    }
    // facet normals at the facet centres
    std::vector<double> N((size_t)3 * nf);
    auto mean_normal = [](const double *c00, const double *c10, const double *c01, const double *c11, double *n) {
        double t1[3], t2[3];      // mean tangents along the first / second in-plane index
        for (int t = 0; t < 3; ++t) { t1[t] = 0.5 * ((c10[t] - c00[t]) + (c11[t] - c01[t])); t2[t] = 0.5 * ((c01[t] - c00[t]) + (c11[t] - c10[t])); }
        n[0] = t1[1] * t2[2] - t1[2] * t2[1]; n[1] = t1[2] * t2[0] - t1[0] * t2[2]; n[2] = t1[0] * t2[1] - t1[1] * t2[0];
    };
    for (int k = 0; k < mesh.nz; ++k) for (int j = 0; j < mesh.ny; ++j) for (int i = 0; i <= mesh.nx; ++i)      // x-faces: (y, z)
        mean_normal(vtx(i, j, k), vtx(i, j + 1, k), vtx(i, j, k + 1), vtx(i, j + 1, k + 1), &N[(size_t)3 * mesh.fx(i, j, k)]);
    for (int k = 0; k < mesh.nz; ++k) for (int j = 0; j <= mesh.ny; ++j) for (int i = 0; i < mesh.nx; ++i)      // y-faces: (z, x)
        mean_normal(vtx(i, j, k), vtx(i, j, k + 1), vtx(i + 1, j, k), vtx(i + 1, j, k + 1), &N[(size_t)3 * mesh.fy(i, j, k)]);
    for (int k = 0; k <= mesh.nz; ++k) for (int j = 0; j < mesh.ny; ++j) for (int i = 0; i < mesh.nx; ++i)      // z-faces: (x, y)
        mean_normal(vtx(i, j, k), vtx(i + 1, j, k), vtx(i, j + 1, k), vtx(i + 1, j + 1, k), &N[(size_t)3 * mesh.fz(i, j, k)]);
    S.facet_area.assign(nf, 0.0);
    for (int f = 0; f < nf; ++f)
    {
        S.facet_area[f] = std::sqrt(N[3 * f] * N[3 * f] + N[3 * f + 1] * N[3 * f + 1] + N[3 * f + 2] * N[3 * f + 2]);
        m21[f] = 1.0 / S.facet_area[f];
    }
    const int ne = mesh.ne();
    std::vector<double> T((size_t)3 * ne, 0.0);     // edge vectors
    if (with_curl)
    {
        // tangential-trace mass matrices of the facets, 2x2 Gauss rule
        double z16[16] = {0};
        fill_pool(S.M[{1, 1}], nf, z16, 4);
        double *m11 = S.M[{1, 1}].vals.data();
        auto facet_mass = [&](const double *c00, const double *c10, const double *c01, const double *c11, double *Mf) {
            for (int q = 0; q < 16; ++q) Mf[q] = 0.0;
            for (int iu = 0; iu < 2; ++iu) for (int iv = 0; iv < 2; ++iv)
            {
                const double u = g2[iu], v = g2[iv];
                double tu[3], tv[3];
                for (int t = 0; t < 3; ++t)
                {
                    tu[t] = (c10[t] - c00[t]) * (1 - v) + (c11[t] - c01[t]) * v;
                    tv[t] = (c01[t] - c00[t]) * (1 - u) + (c11[t] - c10[t]) * u;
                }
                const double a11 = tu[0] * tu[0] + tu[1] * tu[1] + tu[2] * tu[2], a22 = tv[0] * tv[0] + tv[1] * tv[1] + tv[2] * tv[2];
                const double a12 = tu[0] * tv[0] + tu[1] * tv[1] + tu[2] * tv[2], dt = a11 * a22 - a12 * a12;
                // G = J (J^T J)^-1: columns = dual tangents
                double Gu[3], Gv[3];
                for (int t = 0; t < 3; ++t) { Gu[t] = (a22 * tu[t] - a12 * tv[t]) / dt; Gv[t] = (a11 * tv[t] - a12 * tu[t]) / dt; }
                const double wt = w2[iu] * w2[iv] * std::sqrt(dt);
                double Wv[4][3];
                for (int t = 0; t < 3; ++t) { Wv[0][t] = (1 - v) * Gu[t]; Wv[1][t] = v * Gu[t]; Wv[2][t] = (1 - u) * Gv[t]; Wv[3][t] = u * Gv[t]; }
                for (int p = 0; p < 4; ++p) for (int q = 0; q < 4; ++q)
                    Mf[p * 4 + q] += wt * (Wv[p][0] * Wv[q][0] + Wv[p][1] * Wv[q][1] + Wv[p][2] * Wv[q][2]);
            }
        };
        for (int k = 0; k < mesh.nz; ++k) for (int j = 0; j < mesh.ny; ++j) for (int i = 0; i <= mesh.nx; ++i)      // (u,v) = (y,z)
            facet_mass(vtx(i, j, k), vtx(i, j + 1, k), vtx(i, j, k + 1), vtx(i, j + 1, k + 1), m11 + (size_t)16 * mesh.fx(i, j, k));
        for (int k = 0; k < mesh.nz; ++k) for (int j = 0; j <= mesh.ny; ++j) for (int i = 0; i < mesh.nx; ++i)      // (u,v) = (x,z)
            facet_mass(vtx(i, j, k), vtx(i + 1, j, k), vtx(i, j, k + 1), vtx(i + 1, j, k + 1), m11 + (size_t)16 * mesh.fy(i, j, k));
        for (int k = 0; k <= mesh.nz; ++k) for (int j = 0; j < mesh.ny; ++j) for (int i = 0; i < mesh.nx; ++i)      // (u,v) = (x,y)
            facet_mass(vtx(i, j, k), vtx(i + 1, j, k), vtx(i, j + 1, k), vtx(i + 1, j + 1, k), m11 + (size_t)16 * mesh.fz(i, j, k));
        // edges
        auto edge = [&](int id, const double *p0, const double *p1) { for (int t = 0; t < 3; ++t) T[(size_t)3 * id + t] = p1[t] - p0[t]; };
        for (int k = 0; k <= mesh.nz; ++k) for (int j = 0; j <= mesh.ny; ++j) for (int i = 0; i < mesh.nx; ++i) edge(mesh.ex(i, j, k), vtx(i, j, k), vtx(i + 1, j, k));
        for (int k = 0; k <= mesh.nz; ++k) for (int j = 0; j < mesh.ny; ++j) for (int i = 0; i <= mesh.nx; ++i) edge(mesh.ey(i, j, k), vtx(i, j, k), vtx(i, j + 1, k));
        for (int k = 0; k < mesh.nz; ++k) for (int j = 0; j <= mesh.ny; ++j) for (int i = 0; i <= mesh.nx; ++i) edge(mesh.ez(i, j, k), vtx(i, j, k), vtx(i, j, k + 1));
        double one = 1.0;
        fill_pool(S.M[{1, 2}], ne, &one, 1);
        double *m12 = S.M[{1, 2}].vals.data();
        S.ridge_length.assign(ne, 0.0);
        for (int e = 0; e < ne; ++e)
        {
            S.ridge_length[e] = std::sqrt(T[3 * e] * T[3 * e] + T[3 * e + 1] * T[3 * e + 1] + T[3 * e + 2] * T[3 * e + 2]);
            m12[e] = 1.0 / S.ridge_length[e];
        }
    }
    if (with_h1)
    {
        double z16[16] = {0};
        fill_pool(S.M[{0, 1}], nf, z16, 4);
        double *m01 = S.M[{0, 1}].vals.data();
        auto facet_mass_h1 = [&](const double *c00, const double *c10, const double *c01, const double *c11, double *Mf) {
            for (int q = 0; q < 16; ++q) Mf[q] = 0.0;
            for (int iu = 0; iu < 2; ++iu) for (int iv = 0; iv < 2; ++iv)
            {
                const double u = g2[iu], v = g2[iv];
                double tu[3], tv[3];
                for (int t = 0; t < 3; ++t)
                {
                    tu[t] = (c10[t] - c00[t]) * (1 - v) + (c11[t] - c01[t]) * v;
                    tv[t] = (c01[t] - c00[t]) * (1 - u) + (c11[t] - c10[t]) * u;
                }
                const double a11 = tu[0] * tu[0] + tu[1] * tu[1] + tu[2] * tu[2], a22 = tv[0] * tv[0] + tv[1] * tv[1] + tv[2] * tv[2];
                const double a12 = tu[0] * tv[0] + tu[1] * tv[1] + tu[2] * tv[2];
                const double wt = w2[iu] * w2[iv] * std::sqrt(a11 * a22 - a12 * a12);
                const double sh[4] = {(1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v};
                for (int p = 0; p < 4; ++p) for (int q = 0; q < 4; ++q) Mf[p * 4 + q] += wt * sh[p] * sh[q];
            }
        };
        for (int k = 0; k < mesh.nz; ++k) for (int j = 0; j < mesh.ny; ++j) for (int i = 0; i <= mesh.nx; ++i)
            facet_mass_h1(vtx(i, j, k), vtx(i, j + 1, k), vtx(i, j, k + 1), vtx(i, j + 1, k + 1), m01 + (size_t)16 * mesh.fx(i, j, k));
        for (int k = 0; k < mesh.nz; ++k) for (int j = 0; j <= mesh.ny; ++j) for (int i = 0; i < mesh.nx; ++i)
            facet_mass_h1(vtx(i, j, k), vtx(i + 1, j, k), vtx(i, j, k + 1), vtx(i + 1, j, k + 1), m01 + (size_t)16 * mesh.fy(i, j, k));
        for (int k = 0; k <= mesh.nz; ++k) for (int j = 0; j < mesh.ny; ++j) for (int i = 0; i < mesh.nx; ++i)
            facet_mass_h1(vtx(i, j, k), vtx(i + 1, j, k), vtx(i, j + 1, k), vtx(i + 1, j + 1, k), m01 + (size_t)16 * mesh.fz(i, j, k));
        double z4[4] = {0};
        fill_pool(S.M[{0, 2}], ne, z4, 2);
        double *m02 = S.M[{0, 2}].vals.data();
        for (int e = 0; e < ne; ++e)
        {
            const double l = S.ridge_length[e];
            m02[4 * e] = m02[4 * e + 3] = l * (1.0 / 3.0);
            m02[4 * e + 1] = m02[4 * e + 2] = l * (1.0 / 6.0);
        }
        double one = 1.0;
        fill_pool(S.M[{0, 3}], mesh.nv(), &one, 1);
    }
    D.resize(3);
    D[0] = topo->GetB(2);
    D[1] = topo->GetB(1);
    D[2] = topo->GetB(0);
    for (int e = 0; e < nel; ++e)
        for (int q = D[2].I[e]; q < D[2].I[e + 1]; ++q) D[2].A[q] *= (1.0 / vol[e]);
    S.l2const.assign(nel, 1.0);
    if (!with_curl) S.ridge_length.clear();
    S.targets.resize(4); S.ntargets = {4, 3, 3, 1};
    S.targets[3].assign(nel, 1.0);
    S.targets[2].assign((size_t)3 * nf, 0.0);
    for (int f = 0; f < nf; ++f) for (int cc = 0; cc < 3; ++cc) S.targets[2][(size_t)cc * nf + f] = N[3 * f + cc];
    S.targets[1].assign((size_t)3 * ne, 0.0);             // forms below jformStart are not coarsened
    if (with_curl) for (int e = 0; e < ne; ++e) for (int cc = 0; cc < 3; ++cc) S.targets[1][(size_t)cc * ne + e] = T[3 * e + cc];
    S.targets[0].assign((size_t)4 * mesh.nv(), 0.0);
    if (with_h1)
    {
        const size_t nv = (size_t)mesh.nv();
        for (size_t v = 0; v < nv; ++v)
        {
            S.targets[0][v] = 1.0;
            S.targets[0][nv + v] = X[3 * v + 2];
            S.targets[0][2 * nv + v] = X[3 * v + 1];
            S.targets[0][3 * nv + v] = X[3 * v];
        }
    }
}

/// fills SequenceData + D_ for the fine level; alpha / beta: optional per-element weights
/// of the L2 and H(div) element mass matrices (ReplaceMassIntegrator in the drivers)
/// beta_components = 3: beta holds (beta_x, beta_y, beta_z) per element, the diagonal tensor coefficient of the H(div)
/// mass matrix (VectorFunctionCoefficient in VectorFEMassIntegrator, e.g. the SPE10 inverse permeability,
/// examples/MultigridTestSPE10.cpp:377-395); on an axis-aligned cell it scales the three 2 x 2 axis blocks separately.
inline void BuildFineHexSequence(const StructuredHexMesh &mesh, const std::shared_ptr<AgglomeratedTopology> &topo,
                                 const double *alpha, const double *beta, int jstart, SequenceData &S, std::vector<HostCSR> &D,
                                 int beta_components = 1)
{
    using namespace hexfe;
    PARELAG_TEST_FOR_EXCEPTION(beta_components != 1 && beta_components != 3, std::runtime_error, "BuildFineHexSequence: beta has 1 or 3 components per element");
    PARELAG_TEST_FOR_EXCEPTION(mesh.deformed() && beta && beta_components == 3, not_implemented_error,
                               "BuildFineHexSequence: tensor coefficient on trilinear hexahedra");
    if (mesh.deformed()) { BuildFineHexSequenceDeformed(mesh, topo, alpha, beta, jstart, S, D); return; }
    const double hx = mesh.hx, hy = mesh.hy, hz = mesh.hz, vol = hx * hy * hz;
    const int nel = (int)mesh.nel();
    S.topo = topo; S.nforms = 4; S.jstart = jstart; S.is_fe = true;
    S.dof.resize(4);
    Timer t_dofs = TimeManager::AddTimer("Fine sequence: dof handlers");
    for (int j = 0; j < 4; ++j)
    {
        auto dh = std::make_shared<DofHandlerX>(3 - j, topo);
        dh->ndofs = topo->GetNumberLocalEntities(3 - j);
        for (int c = 0; c <= 3 - j; ++c)
            dh->entity_dof[c] = (c == 3 - j) ? hostcsr::Identity(topo->GetNumberLocalEntities(c)) : topo->GetConnectivity(c, 3 - j);
        dh->ComputeBoundaryMask();
        S.dof[j] = dh;
    }
    t_dofs.Stop();
    Timer t_pools = TimeManager::AddTimer("Fine sequence: D and mass pools");
    D.resize(3);
    D[0] = topo->GetB(2);
    D[1] = topo->GetB(1);
    D[2] = topo->GetB(0);
    for (auto &v : D[2].A) v *= (1.0 / vol);
    const double M1[4] = {1.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0, 1.0 / 3.0};
    double K[16];
    kron2(M1, 2, M1, 2, K);
    // form 3
    { double b = vol; fill_pool(S.M[{3, 0}], nel, &b, 1, alpha); }
    // form 2: element (x-,x+,y-,y+,z-,z+), facet
    {
        double blk[36] = {0};
        place(blk, 6, 0, M1, 2, hx / (hy * hz)); place(blk, 6, 2, M1, 2, hy / (hx * hz)); place(blk, 6, 4, M1, 2, hz / (hx * hy));
        fill_pool(S.M[{2, 0}], nel, blk, 6, beta_components == 1 ? beta : nullptr);
        if (beta && beta_components == 3)
        {
            // diagonal tensor: block a (dofs 2a, 2a+1) of element e is scaled by beta[3 e + a]
            double *vals = S.M[{2, 0}].vals.data() + (S.M[{2, 0}].vals.size() - (size_t)nel * 36);
#pragma omp parallel for schedule(static)
            for (int e = 0; e < nel; ++e)
                for (int a = 0; a < 3; ++a)
                    for (int x = 0; x < 2; ++x)
                        for (int y = 0; y < 2; ++y) vals[(size_t)e * 36 + (size_t)(2 * a + x) * 6 + 2 * a + y] *= beta[3 * (size_t)e + a];
        }
        double ax = 1.0 / (hy * hz), ay = 1.0 / (hx * hz), az = 1.0 / (hx * hy);
        fill_pool(S.M[{2, 1}], mesh.nfx(), &ax, 1); fill_pool(S.M[{2, 1}], mesh.nfy(), &ay, 1); fill_pool(S.M[{2, 1}], mesh.nfz(), &az, 1);
    }
    // form 1: element (4 x-edges, 4 y-edges, 4 z-edges), facet (2+2 edges), ridge
    {
        double blk[144] = {0};
        place(blk, 12, 0, K, 4, hy * hz / hx); place(blk, 12, 4, K, 4, hx * hz / hy); place(blk, 12, 8, K, 4, hx * hy / hz);
        fill_pool(S.M[{1, 0}], nel, blk, 12);
        double f[16];
        std::fill(f, f + 16, 0.0); place(f, 4, 0, M1, 2, hz / hy); place(f, 4, 2, M1, 2, hy / hz); fill_pool(S.M[{1, 1}], mesh.nfx(), f, 4);
        std::fill(f, f + 16, 0.0); place(f, 4, 0, M1, 2, hz / hx); place(f, 4, 2, M1, 2, hx / hz); fill_pool(S.M[{1, 1}], mesh.nfy(), f, 4);
        std::fill(f, f + 16, 0.0); place(f, 4, 0, M1, 2, hy / hx); place(f, 4, 2, M1, 2, hx / hy); fill_pool(S.M[{1, 1}], mesh.nfz(), f, 4);
        double lx = 1.0 / hx, ly = 1.0 / hy, lz = 1.0 / hz;
        fill_pool(S.M[{1, 2}], mesh.nex(), &lx, 1); fill_pool(S.M[{1, 2}], mesh.ney(), &ly, 1); fill_pool(S.M[{1, 2}], mesh.nez(), &lz, 1);
    }
    // form 0: element (8 vertices), facet (4), ridge (2), peak
    {
        double K3[64], tmp[64];
        kron2(M1, 2, K, 4, K3);
        for (int q = 0; q < 64; ++q) tmp[q] = vol * K3[q];
        fill_pool(S.M[{0, 0}], nel, tmp, 8);
        double f[16];
        for (int q = 0; q < 16; ++q) f[q] = hy * hz * K[q]; fill_pool(S.M[{0, 1}], mesh.nfx(), f, 4);
        for (int q = 0; q < 16; ++q) f[q] = hx * hz * K[q]; fill_pool(S.M[{0, 1}], mesh.nfy(), f, 4);
        for (int q = 0; q < 16; ++q) f[q] = hx * hy * K[q]; fill_pool(S.M[{0, 1}], mesh.nfz(), f, 4);
        double e[4];
        for (int q = 0; q < 4; ++q) e[q] = hx * M1[q]; fill_pool(S.M[{0, 2}], mesh.nex(), e, 2);
        for (int q = 0; q < 4; ++q) e[q] = hy * M1[q]; fill_pool(S.M[{0, 2}], mesh.ney(), e, 2);
        for (int q = 0; q < 4; ++q) e[q] = hz * M1[q]; fill_pool(S.M[{0, 2}], mesh.nez(), e, 2);
        double one = 1.0;
        fill_pool(S.M[{0, 3}], mesh.nv(), &one, 1);
    }
    t_pools.Stop();
    Timer t_targets = TimeManager::AddTimer("Fine sequence: targets");
    S.l2const.assign(nel, 1.0);
    S.facet_area.assign(mesh.nf(), 0.0);
    std::fill(S.facet_area.begin(), S.facet_area.begin() + mesh.nfx(), hy * hz);
    std::fill(S.facet_area.begin() + mesh.nfx(), S.facet_area.begin() + mesh.nfx() + mesh.nfy(), hx * hz);
    std::fill(S.facet_area.begin() + mesh.nfx() + mesh.nfy(), S.facet_area.end(), hx * hy);
    S.ridge_length.assign(mesh.ne(), 0.0);
    std::fill(S.ridge_length.begin(), S.ridge_length.begin() + mesh.nex(), hx);
    std::fill(S.ridge_length.begin() + mesh.nex(), S.ridge_length.begin() + mesh.nex() + mesh.ney(), hy);
    std::fill(S.ridge_length.begin() + mesh.nex() + mesh.ney(), S.ridge_length.end(), hz);
    // targets (upscaling order 0): L2 {1}; H(div)/H(curl) {e_x,e_y,e_z}; H1 {1,z,y,x}
    S.targets.resize(4); S.ntargets = {4, 3, 3, 1};
    S.targets[3].assign(nel, 1.0);
    {
        const int n = mesh.nf();
        S.targets[2].assign((size_t)3 * n, 0.0);
        for (int f = 0; f < mesh.nfx(); ++f) S.targets[2][f] = S.facet_area[f];
        for (int f = mesh.nfx(); f < mesh.nfx() + mesh.nfy(); ++f) S.targets[2][(size_t)n + f] = S.facet_area[f];
        for (int f = mesh.nfx() + mesh.nfy(); f < n; ++f) S.targets[2][(size_t)2 * n + f] = S.facet_area[f];
    }
    {
        const int n = mesh.ne();
        S.targets[1].assign((size_t)3 * n, 0.0);
        for (int e = 0; e < mesh.nex(); ++e) S.targets[1][e] = S.ridge_length[e];
        for (int e = mesh.nex(); e < mesh.nex() + mesh.ney(); ++e) S.targets[1][(size_t)n + e] = S.ridge_length[e];
        for (int e = mesh.nex() + mesh.ney(); e < n; ++e) S.targets[1][(size_t)2 * n + e] = S.ridge_length[e];
    }
    {
        const int n = mesh.nv();
        S.targets[0].assign((size_t)4 * n, 0.0);
        for (int k = 0; k <= mesh.nz; ++k) for (int j = 0; j <= mesh.ny; ++j) for (int i = 0; i <= mesh.nx; ++i)
        {
            const int v = mesh.vx(i, j, k);
            S.targets[0][v] = 1.0;
            S.targets[0][(size_t)n + v] = mesh.z0 + k * hz;
            S.targets[0][(size_t)2 * n + v] = mesh.y0 + j * hy;
            S.targets[0][(size_t)3 * n + v] = mesh.x0 + i * hx;
        }
    }
}
} // namespace parelag
