"""Oracle for the main/side loop of ``recon.cpp:65-119``.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

from .flow import calculate_flow
from .render import RenderOracle, mix_background
from .tri import triangulate_pixels


def process_main_frame(render: RenderOracle, frames, cameras, fa, sides, use_farneback=False, keep=False):
    """One iteration of the outer loop: depth -> per side (projected, mixBackground,
    calculateFlow) -> triangulatePixels.  ``frames[i]`` are H x W uint8,
    ``cameras[i]`` 4x4 float32.  Returns M x 7 rows (and the intermediates if ``keep``)."""
    original = frames[fa]
    depth = render.depth(cameras[fa])                                   # recon.cpp:70
    inter = dict(depth0=depth.copy(), projected=[], mixed=[], flows=[])
    flows, cams = [], []
    for fb in sides:                                                    # recon.cpp:81
        proj = render.projected(cameras[fa], frames[fb], cameras[fb])   # recon.cpp:85
        mixed = mix_background(proj, original, depth)                   # recon.cpp:86 (mutates depth)
        flow = calculate_flow(original, mixed, use_farneback)           # recon.cpp:89
        flows.append(flow)
        cams.append(cameras[fb])
        if keep:
            inter["projected"].append(proj)
            inter["mixed"].append(mixed)
            inter["flows"].append(flow)
    if keep:
        tri, evals = triangulate_pixels(flows, cameras[fa], cams, depth, return_evals=True)   # recon.cpp:114
        inter["depth"] = depth
        inter["evals"] = evals
        return tri, inter
    tri = triangulate_pixels(flows, cameras[fa], cams, depth)           # recon.cpp:114
    return tri
