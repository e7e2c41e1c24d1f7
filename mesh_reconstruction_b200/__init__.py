"""B200-native dense-correspondence hot path of addam/mesh-reconstruction.

The product is ``libmeshrecon_b200.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/meshrecon_b200.h``).  This package is the thin host-side mirror of the
reference's C++ interface for the path (``recon.hpp:40-55,93-100``): same names, same
argument meaning, same in-place behaviour.  There is no CPU fallback: importing works
anywhere, but every compute call needs the CUDA library and a B200.
"""
from .api import (MeshReconError, Render, calculateFlow, compare, extractCameraCenter, filterPoints, filter_rows, flowRemap,  # noqa: F401
                  imageGradient, library_path, load_library, mixBackground, process_main_frame, spawnRender, submit_main_frame,
                  triangulatePixels)

__all__ = ["MeshReconError", "Render", "calculateFlow", "compare", "extractCameraCenter", "filterPoints", "filter_rows", "flowRemap",
           "imageGradient", "library_path", "load_library", "mixBackground", "process_main_frame", "spawnRender", "submit_main_frame",
           "triangulatePixels"]
