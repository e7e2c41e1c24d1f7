"""CPU oracle for the dense-correspondence hot path of addam/mesh-reconstruction.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``mesh_reconstruction_b200``) never imports, links or executes anything in
here and fails loudly when its CUDA library is missing.

What it restates (file:line are into the upstream reference tree):

* ``flow.cpp:19-42``            -> :mod:`oracle.flow`      (calculateFlow, via the real OpenCV binary ``cv2``)
* ``util.cpp:332-403,465-479``  -> :mod:`oracle.flow` / :mod:`oracle.cvprims` (compare, flowRemap, imageGradient)
* ``util.cpp:33-53,62-329``     -> ``oracle/recon_oracle.c`` + :mod:`oracle.tri` (triangulatePixels / triangulatePixel)
* ``util.cpp:366-387``          -> :mod:`oracle.render`    (mixBackground)
* ``render_glx.cpp:230-397`` + ``shader.vert`` / ``shader.frag`` -> ``oracle/recon_oracle.c`` + :mod:`oracle.render`
* ``recon.cpp:65-119``          -> :mod:`oracle.pipeline`  (the main/side loop)

Parity pinning status
---------------------
The reference's own tests hold no golden vector or known-answer value for this
path, and the reference cannot be built in this image
(OpenCV C++/GLX/CGAL absent), so strictly speaking **parity is unpinned by the
reference's own tests**.  What *is* pinned:

* every OpenCV-owned primitive on the path (VariationalRefinement, remap
  INTER_CUBIC on 8U, pyrDown/pyrUp, Sobel, 4x4 / 2x2 ``invert``, small ``gemm``,
  PCA) is executed by, or checked against, the real OpenCV binary (``cv2`` 4.13)
  in ``tests/test_oracle_cv.py``: VariationalRefinement, remap, pyrDown / pyrUp /
  ``compare``, Sobel, ``invert`` and ``gemm`` restatements are BIT-IDENTICAL to
  cv2 (including its SIMD/scalar column split); PCA eigenvectors agree to float
  rounding where the eigen-gap defines them;
* the camera tracks and bundle points the reference ships (``tracks/*.yaml``) are
  committed as golden inputs (``tests/golden/tracks_*.npz``, made by
  ``tests/golden/make_yaml_fixtures.py``) and drive full-path parity tests;
* the only concrete fixture the reference has on this path (the 25-vertex mesh
  and the two MVP matrices of ``render_glx.cpp:407-410``) is a committed golden
  input (``tests/golden``), with known-answer properties from Appendix E of
  SURVEY.md (all vertices inside the main frustum, NDC z in [0.75, 0.97]).
"""
