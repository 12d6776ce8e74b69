// amge_coarsen.hpp -- entry points of the coarsening path (see amge_coarsen.cpp)
#pragma once
#include "amge_hex.hpp"
#include "amge_tet.hpp"

namespace parelag
{
/// Steps 3-4 of the drivers (examples/MultigridTest2Form.cpp:248-375) on a structured hex
/// mesh: fine topology, nlevels-1 derefinement agglomerations
/// (MFEMRefinedMeshPartitioner + CoarsenLocalPartitioning), fine DeRhamSequence with the
/// order-0 upscaling targets, then Coarsen() level by level.  Timers use the reference's names.
/// vertex_coords (nv x 3, optional, single rank): moved vertices -> trilinear hexahedra (needs jstart >= 2).
/// beta_components = 3: beta is a diagonal tensor (beta_x, beta_y, beta_z) per element (axis-aligned cells).
std::vector<std::shared_ptr<DeRhamSequence>> BuildHexSequenceHierarchy(int nx, int ny, int nz, double Lx, double Ly, double Lz,
                                                                        const double *alpha, const double *beta, int jstart,
                                                                        int nlevels, double svd_tol, const double *vertex_coords = nullptr,
                                                                        int beta_components = 1);
/// The same on one box of a P[0] x P[1] x P[2] box decomposition (one rank <-> one box <-> one GPU,
/// the reference's one-MPI-rank-per-partition model): nx, ny, nz and Lx, Ly, Lz describe THIS
/// rank's box; every rank coarsens its own box (elements never migrate) and the levels are glued
/// by the dof <-> true-dof SharingMaps built at the end (amge_par.hpp).  comm == NULL: serial.
std::vector<std::shared_ptr<DeRhamSequence>> BuildHexSequenceHierarchyPar(const pe_host_comm *comm, const int *procs, int nx, int ny, int nz,
                                                                           double Lx, double Ly, double Lz, const double *alpha,
                                                                           const double *beta, int jstart, int nlevels, double svd_tol,
                                                                           const double *vertex_coords = nullptr, int beta_components = 1);
/// The same on an unstructured tetrahedral mesh (examples/MultigridTest0Form.cpp:147-375, BASELINE configs[0]): the coarse
/// mesh is refined `nref` times (serial + parallel refinements of the driver), the finest mesh is level 0 and the
/// nlevels - 1 <= nref coarser levels come from derefinement (MFEMRefinedMeshPartitioner: partition = element / 8).
std::vector<std::shared_ptr<DeRhamSequence>> BuildTetSequenceHierarchy(const TetMesh &coarse_mesh, int nref, int nlevels, const double *alpha,
                                                                        const double *beta, int jstart, double svd_tol);
} // namespace parelag
