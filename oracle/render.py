"""Oracle for ``Render::{loadMesh,depth,projected}`` (render_glx.cpp:230-397,
shader.vert:9-13, shader.frag:11-25) and ``mixBackground`` (util.cpp:366-387).
TEST INFRASTRUCTURE ONLY.  The GL semantics that the reference leaves to the
driver are DEFINED in ``recon_oracle.c`` (see the comment block there)."""
import numpy as np

from . import native

f32 = np.float32


class RenderOracle:
    """Mirrors ``class Render`` (recon.hpp:93-99)."""

    def __init__(self, width, height):
        self.W, self.H = int(width), int(height)
        self.soup = np.zeros((0, 9), f32)

    def loadMesh(self, vertices, faces):
        v = np.ascontiguousarray(vertices, f32)
        f = np.ascontiguousarray(faces, np.int32)
        self.soup = np.zeros((len(f), 9), f32)
        if len(f):
            native.lib().orc_load_mesh(v, len(v), f, len(f), self.soup)

    def raster(self, camera):
        d = np.empty((self.H, self.W), f32)
        t = np.empty((self.H, self.W), np.int32)
        native.lib().orc_raster(self.soup, len(self.soup), np.ascontiguousarray(camera, f32), self.W, self.H, d, t)
        return d, t

    def depth(self, camera):
        return self.raster(camera)[0]

    def shadow_map(self, projector):
        """Dilated side-camera depth, rows top-down (render_glx.cpp:272-314)."""
        d = self.depth(projector)
        gl = np.ascontiguousarray(d[::-1])
        native.lib().orc_dilate_shadow_gl(gl, self.W, self.H)
        return np.ascontiguousarray(gl[::-1])

    def projected(self, camera, frame, projector):
        assert frame.ndim == 2 and frame.dtype == np.uint8
        out = np.empty((self.H, self.W, 3), np.uint8)
        native.lib().orc_projected(self.soup, len(self.soup), np.ascontiguousarray(camera, f32),
                                   np.ascontiguousarray(frame), np.ascontiguousarray(projector, f32),
                                   self.W, self.H, out)
        return out


def mix_background(image, background, depth):
    """``mixBackground``: returns the mixed 8UC1 image and MUTATES ``depth``."""
    assert image.shape[2] == 3 and background.ndim == 2
    assert depth.dtype == np.float32 and depth.flags.c_contiguous
    H, W = background.shape
    out = np.empty((H, W), np.uint8)
    native.lib().orc_mix_background(np.ascontiguousarray(image), np.ascontiguousarray(background), depth, W, H, out)
    return out


def dilate_shadow_parallel(shadow_td):
    """Closed form of the sequential dilation (used to cross-check the parallel
    formulation the CUDA kernel uses).  Input/outputs rows top-down."""
    s = np.ascontiguousarray(shadow_td[::-1]).astype(f32)  # GL orientation
    H, W = s.shape
    hf = s.copy()
    if W > 2:
        hf[1:, 1:-1] = np.maximum(np.maximum(s[1:, :-2], s[1:, 1:-1]), s[1:, 2:])
        pm = np.minimum.accumulate(s[0])          # prefix min of row 0
        hf[0, 1:-1] = pm[2:]
    out = s.copy()
    if W > 2:
        up = np.vstack([hf[:1], hf[:-1]])         # HF[i-1] (row 0: itself)
        dn = np.vstack([hf[1:], hf[-1:]])         # HF[i+1] (last: itself)
        m = np.maximum(np.maximum(up, hf), dn)
        out[:, 1:-1] = m[:, 1:-1]
    return np.ascontiguousarray(out[::-1])
