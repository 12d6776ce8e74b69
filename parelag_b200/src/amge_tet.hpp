// amge_tet.hpp -- input producer for unstructured tetrahedral meshes: the fine-level de Rham sequence
// (DeRhamSequence3D_FE, src/amge/DeRhamSequenceFE.cpp:633-722, at lowest order) of BASELINE configs[0]
// (examples/MultigridTest0Form.cpp:147-212: meshes/cube456.mesh, uniform refinements, derefinement agglomeration).
//
//   * reader of the NETGEN neutral format of meshes/cube456.mesh (what mfem::Mesh(imesh, 1, 1) accepts there),
//   * uniform (red) refinement; the children of element e are 8e .. 8e+7 -- the numbering
//     MFEMRefinedMeshPartitioner::Partition assumes for MFEM >= 4.1 (partition = element / 8,
//     src/partitioning/MFEMRefinedMeshPartitioner.cpp:48-66),
//   * topology: facets / ridges numbered in lexicographic order of their ascending vertex tuples and oriented by
//     ascending vertex number, boundary facets outward (an mfem::Mesh gives boundary faces the orientation of their
//     element; the topology coarsening groups a boundary patch only if the signs agree, Topology.cpp:735-748),
//   * Whitney forms with MFEM's dof meaning (H1 vertex values, Nedelec edge circulations, Raviart-Thomas face fluxes,
//     L2 cell values, D_2 = net outward flux / volume), element / facet / ridge mass matrices in closed form from the
//     barycentric gradients, PV-trace geometry and the order-0 upscaling targets (SetUpscalingTargets, :927-982).
// Same conventions and arithmetic as oracle/tets.py.
#pragma once
#include <array>
#include <fstream>
#include <sstream>
#include "amge_dofs.hpp"

namespace parelag
{
struct TetMesh
{
    std::vector<double> V;                 // nv x 3
    std::vector<int> T;                    // nel x 4, every row ascending
    std::vector<int> Btri, Battr;          // nb x 3 ascending; attribute (1-based)
    // derived
    std::vector<int> F, E;                 // nf x 3, ne x 2 (lexicographic)
    std::vector<int> el_face, el_edge;     // nel x 4 (012, 013, 023, 123), nel x 6 (01, 02, 03, 12, 13, 23)
    std::vector<int> face_edge;            // nf x 3 (ab, ac, bc)
    std::vector<int> bdr_face;             // nb
    std::vector<double> det, vol, fsign, N, tvec;   // nel, nel, nf, nf x 3 (oriented area vector), ne x 3
    int nv() const { return (int)(V.size() / 3); }
    int nel() const { return (int)(T.size() / 4); }
    int nf() const { return (int)(F.size() / 3); }
    int ne() const { return (int)(E.size() / 2); }
    int nb() const { return (int)Battr.size(); }

    static TetMesh FromArrays(int nv, const double *coords, int nel, const int *tets, int nb, const int *btri, const int *battr)
    {
        TetMesh m;
        m.V.assign(coords, coords + (size_t)3 * nv);
        m.T.assign(tets, tets + (size_t)4 * nel);
        m.Btri.assign(btri, btri + (size_t)3 * nb);
        m.Battr.assign(battr, battr + nb);
        m.Build();
        return m;
    }
    /// NETGEN neutral format: nv, coordinates, ne, (attribute v1 v2 v3 v4), nb, (attribute v1 v2 v3); 1-based vertices
    static TetMesh ReadNetgenNeutral(const std::string &path)
    {
        std::ifstream in(path);
        PARELAG_TEST_FOR_EXCEPTION(!in.good(), std::runtime_error, "Cannot read mesh from input file: " << path);
        std::string tag;
        in >> tag;
        PARELAG_TEST_FOR_EXCEPTION(tag != "NETGEN_Neutral_Format", std::runtime_error, "mesh file " << path << ": not a NETGEN neutral file");
        TetMesh m;
        int nv = 0, ne = 0, nb = 0;
        in >> nv;
        m.V.resize((size_t)3 * nv);
        for (auto &x : m.V) in >> x;
        in >> ne;
        m.T.resize((size_t)4 * ne);
        for (int e = 0; e < ne; ++e) { int a; in >> a; for (int q = 0; q < 4; ++q) { in >> m.T[4 * e + q]; --m.T[4 * e + q]; } }
        in >> nb;
        m.Btri.resize((size_t)3 * nb); m.Battr.resize(nb);
        for (int b = 0; b < nb; ++b) { in >> m.Battr[b]; for (int q = 0; q < 3; ++q) { in >> m.Btri[3 * b + q]; --m.Btri[3 * b + q]; } }
        PARELAG_TEST_FOR_EXCEPTION(in.fail(), std::runtime_error, "mesh file " << path << ": truncated");
        m.Build();
        return m;
    }

    /// MFEM's own format, "MFEM mesh v1.0" (what mfem::Mesh(imesh, 1, 1) reads for the meshes distributed with MFEM):
    /// comment lines start with '#'; sections `dimension`, `elements` (attribute, geometry type, vertices), `boundary`
    /// (attribute, geometry type, vertices), `vertices` (count, space dimension, coordinates); 0-based vertex numbers.
    /// Straight-sided tetrahedral meshes only: geometry types 4 (tetrahedron) for elements and 2 (triangle) for the boundary;
    /// a `nodes` grid function (curved mesh) is rejected.
    static TetMesh ReadMFEM(const std::string &path)
    {
        std::ifstream in(path);
        PARELAG_TEST_FOR_EXCEPTION(!in.good(), std::runtime_error, "Cannot read mesh from input file: " << path);
        // tokens without the comments
        std::vector<std::string> tok;
        bool first = true;
        for (std::string line; std::getline(in, line);)
        {
            if (first)
            {
                first = false;
                PARELAG_TEST_FOR_EXCEPTION(line.rfind("MFEM mesh v1.0", 0) != 0, std::runtime_error, "mesh file " << path << ": not an MFEM mesh v1.0 file");
                continue;
            }
            const size_t hash = line.find('#');
            if (hash != std::string::npos) line.erase(hash);
            std::istringstream ls(line);
            for (std::string t; ls >> t;) tok.push_back(t);
        }
        size_t p = 0;
        auto next = [&]() -> const std::string &
        {
            PARELAG_TEST_FOR_EXCEPTION(p >= tok.size(), std::runtime_error, "mesh file " << path << ": truncated");
            return tok[p++];
        };
        auto expect = [&](const char *word)
        { const std::string &t = next(); PARELAG_TEST_FOR_EXCEPTION(t != word, std::runtime_error, "mesh file " << path << ": expected \"" << word << "\", found \"" << t << "\""); };
        TetMesh m;
        expect("dimension");
        PARELAG_TEST_FOR_EXCEPTION(std::stoi(next()) != 3, std::runtime_error, "mesh file " << path << ": only 3-d meshes");
        expect("elements");
        const int ne = std::stoi(next());
        m.T.resize((size_t)4 * ne);
        for (int e = 0; e < ne; ++e)
        {
            (void)next();                                                  // element attribute
            PARELAG_TEST_FOR_EXCEPTION(std::stoi(next()) != 4, not_implemented_error, "mesh file " << path << ": only tetrahedra (geometry type 4)");
            for (int q = 0; q < 4; ++q) m.T[4 * (size_t)e + q] = std::stoi(next());
        }
        expect("boundary");
        const int nb = std::stoi(next());
        m.Btri.resize((size_t)3 * nb); m.Battr.resize(nb);
        for (int b = 0; b < nb; ++b)
        {
            m.Battr[b] = std::stoi(next());
            PARELAG_TEST_FOR_EXCEPTION(std::stoi(next()) != 2, not_implemented_error, "mesh file " << path << ": only triangular boundary elements (geometry type 2)");
            for (int q = 0; q < 3; ++q) m.Btri[3 * (size_t)b + q] = std::stoi(next());
        }
        expect("vertices");
        const int nv = std::stoi(next());
        const std::string &vd = next();
        PARELAG_TEST_FOR_EXCEPTION(vd == "nodes", not_implemented_error, "mesh file " << path << ": curved meshes (nodes grid function) are not supported");
        PARELAG_TEST_FOR_EXCEPTION(std::stoi(vd) != 3, std::runtime_error, "mesh file " << path << ": vertices must have 3 coordinates");
        m.V.resize((size_t)3 * nv);
        for (auto &x : m.V) x = std::stod(next());
        m.Build();
        return m;
    }
    /// mfem::Mesh(imesh, 1, 1): the format is recognised from the first line
    static TetMesh Read(const std::string &path)
    {
        std::ifstream in(path);
        PARELAG_TEST_FOR_EXCEPTION(!in.good(), std::runtime_error, "Cannot read mesh from input file: " << path);
        std::string line;
        std::getline(in, line);
        in.close();
        if (line.rfind("MFEM mesh v1.0", 0) == 0) return ReadMFEM(path);
        return ReadNetgenNeutral(path);
    }

    void Build()
    {
        const int n = nel();
        for (int e = 0; e < n; ++e) std::sort(T.begin() + 4 * e, T.begin() + 4 * e + 4);
        for (int b = 0; b < nb(); ++b) std::sort(Btri.begin() + 3 * b, Btri.begin() + 3 * b + 3);
        // faces
        {
            static const int fc[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
            std::vector<std::array<int, 3>> all((size_t)4 * n);
#pragma omp parallel for schedule(static)
            for (int e = 0; e < n; ++e)
                for (int q = 0; q < 4; ++q) all[(size_t)4 * e + q] = {T[4 * e + fc[q][0]], T[4 * e + fc[q][1]], T[4 * e + fc[q][2]]};
            std::vector<std::array<int, 3>> uniq(all);
            std::sort(uniq.begin(), uniq.end());
            uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
            F.resize(3 * uniq.size());
            for (size_t f = 0; f < uniq.size(); ++f) { F[3 * f] = uniq[f][0]; F[3 * f + 1] = uniq[f][1]; F[3 * f + 2] = uniq[f][2]; }
            el_face.resize((size_t)4 * n);
#pragma omp parallel for schedule(static)
            for (int64_t k = 0; k < (int64_t)4 * n; ++k)
                el_face[k] = (int)(std::lower_bound(uniq.begin(), uniq.end(), all[k]) - uniq.begin());
            bdr_face.resize(nb());
            for (int b = 0; b < nb(); ++b)
            {
                const std::array<int, 3> key = {Btri[3 * b], Btri[3 * b + 1], Btri[3 * b + 2]};
                auto it = std::lower_bound(uniq.begin(), uniq.end(), key);
                PARELAG_TEST_FOR_EXCEPTION(it == uniq.end() || *it != key, std::runtime_error, "TetMesh: boundary triangle " << b << " is not a face of the mesh");
                bdr_face[b] = (int)(it - uniq.begin());
            }
        }
        // edges
        {
            static const int ed[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
            std::vector<std::array<int, 2>> all((size_t)6 * n);
#pragma omp parallel for schedule(static)
            for (int e = 0; e < n; ++e)
                for (int q = 0; q < 6; ++q) all[(size_t)6 * e + q] = {T[4 * e + ed[q][0]], T[4 * e + ed[q][1]]};
            std::vector<std::array<int, 2>> uniq(all);
            std::sort(uniq.begin(), uniq.end());
            uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
            E.resize(2 * uniq.size());
            for (size_t k = 0; k < uniq.size(); ++k) { E[2 * k] = uniq[k][0]; E[2 * k + 1] = uniq[k][1]; }
            el_edge.resize((size_t)6 * n);
#pragma omp parallel for schedule(static)
            for (int64_t k = 0; k < (int64_t)6 * n; ++k)
                el_edge[k] = (int)(std::lower_bound(uniq.begin(), uniq.end(), all[k]) - uniq.begin());
            face_edge.resize((size_t)3 * nf());
#pragma omp parallel for schedule(static)
            for (int f = 0; f < nf(); ++f)
            {
                const int a = F[3 * f], b = F[3 * f + 1], c = F[3 * f + 2];
                const std::array<int, 2> k0 = {a, b}, k1 = {a, c}, k2 = {b, c};
                face_edge[3 * f] = (int)(std::lower_bound(uniq.begin(), uniq.end(), k0) - uniq.begin());
                face_edge[3 * f + 1] = (int)(std::lower_bound(uniq.begin(), uniq.end(), k1) - uniq.begin());
                face_edge[3 * f + 2] = (int)(std::lower_bound(uniq.begin(), uniq.end(), k2) - uniq.begin());
            }
        }
        // geometry
        det.resize(n); vol.resize(n);
#pragma omp parallel for schedule(static)
        for (int e = 0; e < n; ++e)
        {
            const double *x0 = &V[3 * T[4 * e]], *x1 = &V[3 * T[4 * e + 1]], *x2 = &V[3 * T[4 * e + 2]], *x3 = &V[3 * T[4 * e + 3]];
            double a[3], b[3], c[3];
            for (int k = 0; k < 3; ++k) { a[k] = x1[k] - x0[k]; b[k] = x2[k] - x0[k]; c[k] = x3[k] - x0[k]; }
            det[e] = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
            vol[e] = std::fabs(det[e]) / 6.0;
        }
        // facet orientation: ascending vertex order, boundary facets outward
        fsign.assign(nf(), 1.0);
        {
            std::vector<int> cnt(nf(), 0);
            for (int k : el_face) cnt[k]++;
            for (int e = 0; e < n; ++e)
            {
                const double s = det[e] > 0 ? 1.0 : -1.0;
                const double out[4] = {-s, s, -s, s};            // outward sign of (012), (013), (023), (123)
                for (int q = 0; q < 4; ++q) { const int f = el_face[4 * e + q]; if (cnt[f] == 1) fsign[f] = out[q]; }
            }
        }
        N.resize((size_t)3 * nf());
#pragma omp parallel for schedule(static)
        for (int f = 0; f < nf(); ++f)
        {
            const double *a = &V[3 * F[3 * f]], *b = &V[3 * F[3 * f + 1]], *c = &V[3 * F[3 * f + 2]];
            const double u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, w[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
            N[3 * f] = 0.5 * (u[1] * w[2] - u[2] * w[1]) * fsign[f];
            N[3 * f + 1] = 0.5 * (u[2] * w[0] - u[0] * w[2]) * fsign[f];
            N[3 * f + 2] = 0.5 * (u[0] * w[1] - u[1] * w[0]) * fsign[f];
        }
        tvec.resize((size_t)3 * ne());
        for (int k = 0; k < ne(); ++k) for (int q = 0; q < 3; ++q) tvec[3 * k + q] = V[3 * E[2 * k + 1] + q] - V[3 * E[2 * k] + q];
    }

    /// uniform refinement; child c of element e is element 8e + c, the midpoint of edge q is vertex nv + q
    TetMesh Refine() const
    {
        TetMesh m;
        const int n = nel(), nv0 = nv();
        m.V = V;
        m.V.resize((size_t)3 * (nv0 + ne()));
        for (int k = 0; k < ne(); ++k) for (int q = 0; q < 3; ++q) m.V[3 * (size_t)(nv0 + k) + q] = 0.5 * (V[3 * E[2 * k] + q] + V[3 * E[2 * k + 1] + q]);
        m.T.resize((size_t)32 * n);
#pragma omp parallel for schedule(static)
        for (int e = 0; e < n; ++e)
        {
            const int v0 = T[4 * e], v1 = T[4 * e + 1], v2 = T[4 * e + 2], v3 = T[4 * e + 3];
            const int *ee = &el_edge[6 * e];
            const int m01 = nv0 + ee[0], m02 = nv0 + ee[1], m03 = nv0 + ee[2], m12 = nv0 + ee[3], m13 = nv0 + ee[4], m23 = nv0 + ee[5];
            const int kids[8][4] = {{v0, m01, m02, m03}, {m01, v1, m12, m13}, {m02, m12, v2, m23}, {m03, m13, m23, v3},
                                    {m01, m02, m03, m13}, {m01, m02, m12, m13}, {m02, m03, m13, m23}, {m02, m12, m13, m23}};
            for (int c = 0; c < 8; ++c) for (int q = 0; q < 4; ++q) m.T[(size_t)32 * e + 4 * c + q] = kids[c][q];
        }
        m.Btri.resize((size_t)12 * nb()); m.Battr.resize((size_t)4 * nb());
        for (int b = 0; b < nb(); ++b)
        {
            const int a = Btri[3 * b], bb = Btri[3 * b + 1], c = Btri[3 * b + 2];
            const int *fe = &face_edge[3 * bdr_face[b]];
            const int mab = nv0 + fe[0], mac = nv0 + fe[1], mbc = nv0 + fe[2];
            const int kids[4][3] = {{a, mab, mac}, {mab, bb, mbc}, {mac, mbc, c}, {mab, mac, mbc}};
            for (int k = 0; k < 4; ++k) { for (int q = 0; q < 3; ++q) m.Btri[(size_t)12 * b + 3 * k + q] = kids[k][q]; m.Battr[(size_t)4 * b + k] = Battr[b]; }
        }
        m.Build();
        return m;
    }

    std::shared_ptr<AgglomeratedTopology> Topology() const
    {
        HostCSR B0, B1, B2, fb;
        const int n = nel();
        B0.nrows = n; B0.ncols = nf(); B0.I.resize((size_t)n + 1); B0.J.resize((size_t)4 * n); B0.A.resize((size_t)4 * n);
        for (int e = 0; e <= n; ++e) B0.I[e] = 4 * e;
#pragma omp parallel for schedule(static)
        for (int e = 0; e < n; ++e)
        {
            const double s = det[e] > 0 ? 1.0 : -1.0;
            const double out[4] = {-s, s, -s, s};
            for (int q = 0; q < 4; ++q) { const int f = el_face[4 * e + q]; B0.J[4 * e + q] = f; B0.A[4 * e + q] = out[q] * fsign[f]; }   // ascending
        }
        B1.nrows = nf(); B1.ncols = ne(); B1.I.resize((size_t)nf() + 1); B1.J.resize((size_t)3 * nf()); B1.A.resize((size_t)3 * nf());
        for (int f = 0; f <= nf(); ++f) B1.I[f] = 3 * f;
        for (int f = 0; f < nf(); ++f)
        {
            const double sg[3] = {1.0, -1.0, 1.0};              // boundary of (a b c) = ab - ac + bc
            for (int q = 0; q < 3; ++q) { B1.J[3 * f + q] = face_edge[3 * f + q]; B1.A[3 * f + q] = sg[q] * fsign[f]; }
        }
        B2.nrows = ne(); B2.ncols = nv(); B2.I.resize((size_t)ne() + 1); B2.J.resize((size_t)2 * ne()); B2.A.resize((size_t)2 * ne());
        for (int k = 0; k <= ne(); ++k) B2.I[k] = 2 * k;
        for (int k = 0; k < ne(); ++k) { B2.J[2 * k] = E[2 * k]; B2.A[2 * k] = -1.0; B2.J[2 * k + 1] = E[2 * k + 1]; B2.A[2 * k + 1] = 1.0; }
        int nattr = 1;
        for (int a : Battr) nattr = std::max(nattr, a);
        std::vector<int> attr(nf(), -1);
        for (int b = 0; b < nb(); ++b) attr[bdr_face[b]] = Battr[b] - 1;
        fb.nrows = nf(); fb.ncols = nattr; fb.I = {0};
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

namespace tetfe
{
inline void cross(const double *a, const double *b, double *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/// one pool with `count` blocks of order m, filled in parallel by fn(e, block)
template <class Fn>
inline void fill_blocks(BlockPool &P, int count, int m, Fn fn)
{
    const size_t mm = (size_t)m * m;
    P.size.assign((size_t)count, m);
    P.off.resize((size_t)count + 1);
    P.vals.resize((size_t)count * mm);
    for (int e = 0; e <= count; ++e) P.off[e] = (int64_t)((size_t)e * mm);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < count; ++e) fn(e, P.vals.data() + (size_t)e * mm);
}
/// Whitney 1-form mass block from the gradients g (nvtx x 3) and I_ab = measure (1 + delta_ab) / denom
inline void nedelec_block(const double (*g)[3], int nvtx, const int (*ed)[2], int ned, double measure, double denom, double *M)
{
    auto I = [&](int a, int b) { return measure * (a == b ? 2.0 : 1.0) / denom; };
    for (int p = 0; p < ned; ++p)
        for (int q = 0; q < ned; ++q)
        {
            const int i = ed[p][0], j = ed[p][1], k = ed[q][0], l = ed[q][1];
            M[p * ned + q] = I(i, k) * dot3(g[j], g[l]) - I(i, l) * dot3(g[j], g[k]) - I(j, k) * dot3(g[i], g[l]) + I(j, l) * dot3(g[i], g[k]);
        }
    (void)nvtx;
}
} // namespace tetfe

/// fills SequenceData + D_ for the fine level on tetrahedra; alpha / beta: optional per-element weights of the L2 and
/// H(div) element mass matrices
inline void BuildFineTetSequence(const TetMesh &mesh, const std::shared_ptr<AgglomeratedTopology> &topo, const double *alpha,
                                 const double *beta, int jstart, SequenceData &S, std::vector<HostCSR> &D)
{
    using namespace tetfe;
    const int nel = mesh.nel(), nf = mesh.nf(), ne = mesh.ne(), nv = mesh.nv();
    S.topo = topo; S.nforms = 4; S.jstart = jstart; S.is_fe = true;
    S.dof.resize(4);
    {
        Timer t = TimeManager::AddTimer("Fine sequence: dof handlers");
        for (int j = 0; j < 4; ++j)
        {
            auto dh = std::make_shared<DofHandlerX>(3 - j, topo);
            dh->ndofs = topo->GetNumberLocalEntities(3 - j);
            for (int c = 0; c <= 3 - j; ++c)
                dh->entity_dof[c] = (c == 3 - j) ? hostcsr::Identity(topo->GetNumberLocalEntities(c)) : topo->GetConnectivity(c, 3 - j);
            dh->ComputeBoundaryMask();
            S.dof[j] = dh;
        }
    }
    Timer t_pools = TimeManager::AddTimer("Fine sequence: D and mass pools");
    D.resize(3);
    D[0] = topo->GetB(2);
    D[1] = topo->GetB(1);
    D[2] = topo->GetB(0);
    for (int e = 0; e < nel; ++e)
        for (int q = D[2].I[e]; q < D[2].I[e + 1]; ++q) D[2].A[q] *= (1.0 / mesh.vol[e]);
    const bool with_curl = jstart <= 1, with_h1 = jstart <= 0;
    // gradients of the barycentric coordinates of every element
    std::vector<double> G((size_t)12 * nel);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nel; ++e)
    {
        const double *x0 = &mesh.V[3 * mesh.T[4 * e]], *x1 = &mesh.V[3 * mesh.T[4 * e + 1]], *x2 = &mesh.V[3 * mesh.T[4 * e + 2]],
                     *x3 = &mesh.V[3 * mesh.T[4 * e + 3]];
        double a[3], b[3], c[3], g1[3], g2[3], g3[3];
        for (int k = 0; k < 3; ++k) { a[k] = x1[k] - x0[k]; b[k] = x2[k] - x0[k]; c[k] = x3[k] - x0[k]; }
        cross(b, c, g1); cross(c, a, g2); cross(a, b, g3);
        const double id = 1.0 / mesh.det[e];
        double *g = &G[(size_t)12 * e];
        for (int k = 0; k < 3; ++k) { g[3 + k] = g1[k] * id; g[6 + k] = g2[k] * id; g[9 + k] = g3[k] * id; g[k] = -(g[3 + k] + g[6 + k] + g[9 + k]); }
    }
    fill_blocks(S.M[{3, 0}], nel, 1, [&](int e, double *M) { M[0] = mesh.vol[e] * (alpha ? alpha[e] : 1.0); });
    static const int fc[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
    static const int ed[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    static const int ed3[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    fill_blocks(S.M[{2, 0}], nel, 4, [&](int e, double *M)
    {
        // w_f = 2 (l_a g_b x g_c + l_b g_c x g_a + l_c g_a x g_b); C[p][v] = coefficient vector of l_v in w_p
        const double(*g)[3] = reinterpret_cast<const double(*)[3]>(&G[(size_t)12 * e]);
        double C[4][4][3] = {};
        for (int p = 0; p < 4; ++p)
        {
            const int a = fc[p][0], b = fc[p][1], c = fc[p][2];
            cross(g[b], g[c], C[p][a]); cross(g[c], g[a], C[p][b]); cross(g[a], g[b], C[p][c]);
            for (int v = 0; v < 4; ++v) for (int k = 0; k < 3; ++k) C[p][v][k] *= 2.0;
        }
        const double w = beta ? beta[e] : 1.0;
        for (int p = 0; p < 4; ++p)
            for (int q = 0; q < 4; ++q)
            {
                double s = 0.0;
                for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) s += dot3(C[p][a], C[q][b]) * (mesh.vol[e] * (a == b ? 2.0 : 1.0) / 20.0);
                M[p * 4 + q] = w * s * mesh.fsign[mesh.el_face[4 * e + p]] * mesh.fsign[mesh.el_face[4 * e + q]];
            }
    });
    S.facet_area.assign(nf, 0.0);
    for (int f = 0; f < nf; ++f) S.facet_area[f] = std::sqrt(dot3(&mesh.N[3 * f], &mesh.N[3 * f]));
    S.ridge_length.assign(ne, 0.0);
    for (int k = 0; k < ne; ++k) S.ridge_length[k] = std::sqrt(dot3(&mesh.tvec[3 * k], &mesh.tvec[3 * k]));
    fill_blocks(S.M[{2, 1}], nf, 1, [&](int f, double *M) { M[0] = 1.0 / S.facet_area[f]; });
    if (with_curl)
    {
        fill_blocks(S.M[{1, 0}], nel, 6, [&](int e, double *M)
        { nedelec_block(reinterpret_cast<const double(*)[3]>(&G[(size_t)12 * e]), 4, ed, 6, mesh.vol[e], 20.0, M); });
        fill_blocks(S.M[{1, 1}], nf, 3, [&](int f, double *M)
        {
            // surface gradients of the barycentric coordinates of the triangle
            const double *a = &mesh.V[3 * mesh.F[3 * f]], *b = &mesh.V[3 * mesh.F[3 * f + 1]], *c = &mesh.V[3 * mesh.F[3 * f + 2]];
            const double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
            const double g11 = dot3(e1, e1), g12 = dot3(e1, e2), g22 = dot3(e2, e2), dt = g11 * g22 - g12 * g12;
            double g[3][3];
            for (int k = 0; k < 3; ++k)
            {
                g[1][k] = (g22 * e1[k] - g12 * e2[k]) / dt;
                g[2][k] = (g11 * e2[k] - g12 * e1[k]) / dt;
                g[0][k] = -g[1][k] - g[2][k];
            }
            nedelec_block(g, 3, ed3, 3, S.facet_area[f], 12.0, M);
        });
        fill_blocks(S.M[{1, 2}], ne, 1, [&](int k, double *M) { M[0] = 1.0 / S.ridge_length[k]; });
    }
    if (with_h1)
    {
        fill_blocks(S.M[{0, 0}], nel, 4, [&](int e, double *M) { for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) M[a * 4 + b] = mesh.vol[e] * (a == b ? 2.0 : 1.0) / 20.0; });
        fill_blocks(S.M[{0, 1}], nf, 3, [&](int f, double *M) { for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) M[a * 3 + b] = S.facet_area[f] * (a == b ? 2.0 : 1.0) / 12.0; });
        fill_blocks(S.M[{0, 2}], ne, 2, [&](int k, double *M) { for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) M[a * 2 + b] = S.ridge_length[k] * (a == b ? 2.0 : 1.0) / 6.0; });
        fill_blocks(S.M[{0, 3}], nv, 1, [&](int, double *M) { M[0] = 1.0; });
    }
    t_pools.Stop();
    Timer t_targets = TimeManager::AddTimer("Fine sequence: targets");
    S.l2const.assign(nel, 1.0);
    if (!with_curl) S.ridge_length.clear();
    S.targets.resize(4); S.ntargets = {4, 3, 3, 1};
    S.targets[3].assign(nel, 1.0);
    S.targets[2].assign((size_t)3 * nf, 0.0);
    for (int f = 0; f < nf; ++f) for (int cc = 0; cc < 3; ++cc) S.targets[2][(size_t)cc * nf + f] = mesh.N[3 * f + cc];
    S.targets[1].assign((size_t)3 * ne, 0.0);
    if (with_curl) for (int k = 0; k < ne; ++k) for (int cc = 0; cc < 3; ++cc) S.targets[1][(size_t)cc * ne + k] = mesh.tvec[3 * k + cc];
    S.targets[0].assign((size_t)4 * nv, 0.0);
    if (with_h1)
        for (size_t v = 0; v < (size_t)nv; ++v)
        {
            S.targets[0][v] = 1.0;
            S.targets[0][(size_t)nv + v] = mesh.V[3 * v + 2];
            S.targets[0][(size_t)2 * nv + v] = mesh.V[3 * v + 1];
            S.targets[0][(size_t)3 * nv + v] = mesh.V[3 * v];
        }
}
} // namespace parelag
