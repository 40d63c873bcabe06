"""CPU oracle for the fesr_b200 hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy / CPU PyTorch, the algorithms of the reference
(cmudrc/fast-eng-super-resolution) on the path SURVEY.md section 8 scopes:
graph build, subdomain decomposition, the two mesh models, the node-weight
reduction, overlap stitching and ALDS routing.  Each function cites the
reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it -- as the checker or as the timed CPU
baseline, never as part of the product path (``fesr_b200`` never imports
``oracle`` and fails loudly when ``libfesr.so`` is missing).

Parity pinning (see tests/golden/README.md and DESIGN.md section 3):
  * models.py / GradientbasedLoss / vtk_to_pyg / PCA+KMeans routing are PINNED
    against the reference's own classes executed in the build container
    (tests/golden/make_golden.py imports them from /root/reference with
    torch_geometric.MessagePassing and vtk replaced by stubs) -- fixtures are
    committed under tests/golden/.
  * kd partition and overlap stitch are PARITY UNPINNED: the reference delegates
    them to VTK 9.4.1 C++ filters (vtkRedistributeDataSetFilter,
    vtkStaticPointLocator) that are neither vendored nor installable here and
    ships no golden outputs; the oracle defines them (exact-median kd bisection;
    mean over coincident points keyed by global node id).
  * the two steps either side of the path are PARITY UNPINNED for the same reason: the
    Gaussian-kernel point interpolation (graph.interp_gaussian: vtkPointInterpolator + vtkGaussianKernel) and the
    wall-shear-stress chain (graph.wall_shear_stress: vtkGradientFilter, vtkDataSetSurfaceFilter,
    vtkPolyDataNormals) are restated from the published algorithms of those filters; where the restatement
    deliberately differs (outward face orientation, no feature-edge splitting) the docstring says so.
"""
