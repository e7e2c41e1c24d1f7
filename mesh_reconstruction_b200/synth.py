"""Synthetic input sequences for the BASELINE.json configs (host-side input
synthesis; not part of the hot path).

The reference's three clips are missing upstream (``.MISSING_LARGE_BLOBS:1-3``),
so every benchmark / parity input is synthesised: an analytic height field
``z = amp * sin(fx*x) * sin(fy*y)`` carrying an analytic band-limited texture,
ray-cast from cameras that follow the reference's exporter convention
(``io_export_tracks.py:22-28,59-66``):

    projection = PerspectiveMatrix(fov, aspect, near, far) * inverse(cam * flipZ)

with the 4x4 row-major float32 matrix acting on column vectors, rows giving clip
x, y, z, w.  The proxy mesh handed to ``loadMesh`` is a
coarse triangulation of a *perturbed* height field, so that the reprojected side
frame differs from the true main frame by a sub-pixel residual, which is what the
reference's variational refinement is asked to recover.
"""
from dataclasses import dataclass

import numpy as np

f32 = np.float32


def perspective_matrix(fovx, aspect, near, far):
    """``PerspectiveMatrix`` of io_export_tracks.py:22-28 (float64)."""
    return np.array([[2 / fovx, 0, 0, 0],
                     [0, 2 * aspect / fovx, 0, 0],
                     [0, 0, (far + near) / (far - near), (2 * far * near) / (near - far)],
                     [0, 0, 1, 0]], np.float64)


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """Camera-to-world matrix of a camera at ``eye`` looking along its local +z
    (i.e. Blender's camera matrix already multiplied by flipZ)."""
    eye = np.asarray(eye, np.float64)
    fwd = np.asarray(target, np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, np.float64))
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    c = np.eye(4)
    c[:3, 0], c[:3, 1], c[:3, 2], c[:3, 3] = right, upv, fwd, eye
    return c


@dataclass
class Scene:
    width: int
    height: int
    cameras: np.ndarray      # n x 4 x 4 float32 projection matrices
    cam2world: np.ndarray    # n x 4 x 4 float64
    vertices: np.ndarray     # V x 4 float32 homogeneous (proxy mesh)
    faces: np.ndarray        # F x 3 int32
    amp: float
    fx: float
    fy: float
    tex: np.ndarray          # K x 4 (kx, ky, phase, amplitude)
    fov: float
    near: float
    far: float
    scale: float             # scene scale (bbox diagonal of the proxy mesh)

    def height_at(self, x, y):
        return self.amp * np.sin(self.fx * x) * np.sin(self.fy * y)

    def texture_at(self, x, y):
        v = np.full(np.shape(x), 128.0)
        for kx, ky, ph, a in self.tex:
            v = v + a * np.sin(kx * x + ky * y + ph)
        return v

    def frame(self, i):
        """Ray-cast frame ``i`` -> H x W uint8 (top-down rows)."""
        W, H = self.width, self.height
        c = self.cam2world[i]
        aspect = W / H
        xs = ((np.arange(W) + 0.5) * 2.0 / W - 1.0) * (self.fov / 2)
        ys = (1.0 - (np.arange(H) + 0.5) * 2.0 / H) * (self.fov / (2 * aspect))
        dx, dy = np.meshgrid(xs, ys)
        d = dx[..., None] * c[:3, 0] + dy[..., None] * c[:3, 1] + c[:3, 2]
        o = c[:3, 3]
        t = -o[2] / d[..., 2]
        for _ in range(6):  # Newton on z(t) - h(x(t), y(t)) = 0
            x = o[0] + t * d[..., 0]
            y = o[1] + t * d[..., 1]
            h = self.height_at(x, y)
            hx = self.amp * self.fx * np.cos(self.fx * x) * np.sin(self.fy * y)
            hy = self.amp * self.fy * np.sin(self.fx * x) * np.cos(self.fy * y)
            g = o[2] + t * d[..., 2] - h
            t = t - g / (d[..., 2] - hx * d[..., 0] - hy * d[..., 1])
        x = o[0] + t * d[..., 0]
        y = o[1] + t * d[..., 1]
        return np.clip(np.rint(self.texture_at(x, y)), 0, 255).astype(np.uint8)

    def frame_torch(self, i, device="cuda"):
        """Same ray-cast as :meth:`frame`, evaluated with torch on the GPU (float64).  Used by
        bench.py to synthesise long 1080p / 4K sequences quickly; not bit-identical to the
        NumPy version (different libm), which is irrelevant for throughput inputs."""
        import torch
        W, H = self.width, self.height
        dt = torch.float64
        c = torch.as_tensor(self.cam2world[i], dtype=dt, device=device)
        aspect = W / H
        xs = ((torch.arange(W, dtype=dt, device=device) + 0.5) * 2.0 / W - 1.0) * (self.fov / 2)
        ys = (1.0 - (torch.arange(H, dtype=dt, device=device) + 0.5) * 2.0 / H) * (self.fov / (2 * aspect))
        dy, dx = torch.meshgrid(ys, xs, indexing="ij")
        d = dx[..., None] * c[:3, 0] + dy[..., None] * c[:3, 1] + c[:3, 2]
        o = c[:3, 3]
        t = -o[2] / d[..., 2]
        for _ in range(6):
            x = o[0] + t * d[..., 0]
            y = o[1] + t * d[..., 1]
            sx, cx, sy, cy = torch.sin(self.fx * x), torch.cos(self.fx * x), torch.sin(self.fy * y), torch.cos(self.fy * y)
            g = o[2] + t * d[..., 2] - self.amp * sx * sy
            t = t - g / (d[..., 2] - self.amp * self.fx * cx * sy * d[..., 0] - self.amp * self.fy * sx * cy * d[..., 1])
        x = o[0] + t * d[..., 0]
        y = o[1] + t * d[..., 1]
        v = torch.full_like(x, 128.0)
        for kx, ky, ph, a in self.tex:
            v = v + a * torch.sin(kx * x + ky * y + ph)
        return torch.clamp(torch.round(v), 0, 255).to(torch.uint8)

    def frames(self, idx=None):
        idx = range(len(self.cameras)) if idx is None else idx
        return [self.frame(i) for i in idx]


def make_scene(width, height, n_frames, seed=0, step=0.01, mesh_res=24, mesh_err=0.004,
               extent=1.6, amp=0.05, dist=3.0, fov=0.92, tex_components=24, z0=0.0):
    """Build the synthetic sequence of BASELINE config 4/5 at any size.

    The camera travels on an arc above the surface, ``step`` world units per
    frame (adjacent-frame image motion of the *residual* stays sub-pixel because
    the proxy mesh explains most of it).  ``mesh_err`` is the amplitude of the
    deliberate height error of the proxy mesh."""
    rng = np.random.default_rng(seed)
    fxs, fys = 2.3, 1.9
    tex = np.stack([rng.uniform(-18, 18, tex_components), rng.uniform(-18, 18, tex_components),
                    rng.uniform(0, 2 * np.pi, tex_components), rng.uniform(6, 16, tex_components)], 1)
    # proxy mesh: regular grid over [-extent, extent]^2, perturbed heights
    g = np.linspace(-extent, extent, mesh_res + 1)
    gx, gy = np.meshgrid(g, g)
    gz = amp * np.sin(fxs * gx) * np.sin(fys * gy) + mesh_err * np.sin(5.1 * gx + 0.3) * np.cos(4.3 * gy - 0.2)
    verts = np.stack([gx.ravel(), gy.ravel(), gz.ravel(), np.ones(gx.size)], 1).astype(f32)
    faces = []
    n = mesh_res + 1
    for j in range(mesh_res):
        for i in range(mesh_res):
            a, b, c, d = j * n + i, j * n + i + 1, (j + 1) * n + i, (j + 1) * n + i + 1
            faces += [(a, b, d), (a, d, c)]
    faces = np.asarray(faces, np.int32)
    near, far = 0.5 * dist, 2.0 * dist
    persp = perspective_matrix(fov, width / height, near, far)
    rng_cam = np.random.default_rng(seed + 1)
    phase = rng_cam.uniform(0, 0.2)
    cams, c2w = [], []
    for i in range(n_frames):
        s = (i - (n_frames - 1) / 2) * step
        eye = np.array([s, -0.35 * dist + 0.2 * s + phase * 0.0, dist * 0.9])
        c = look_at(eye, (0.15 * s, 0.0, 0.0))
        c2w.append(c)
        cams.append((persp @ np.linalg.inv(c)).astype(f32))
    if z0 != 0.0:
        T = np.eye(4)
        T[2, 3] = z0                                        # world = T . scene
        Ti = np.linalg.inv(T)
        verts = (verts.astype(np.float64) @ T.T).astype(f32)
        cams = [(persp @ np.linalg.inv(c) @ Ti).astype(f32) for c in c2w]
    lo, hi = verts[:, :3].min(0), verts[:, :3].max(0)
    return Scene(width, height, np.stack(cams), np.stack(c2w), verts, faces, amp, fxs, fys, tex,
                 fov, near, far, float(np.linalg.norm(hi - lo)))


# The only concrete fixture the reference holds on this path: the mesh and the two
# matrices hard-coded in its GL smoke test (render_glx.cpp:407-410).  Known-input
# (no known-answer upstream); values are data, reproduced for the golden tests.
TEST_GLX_POINTS = np.array([
    0.5127, -3.9222, -29.4300, 1.0, 0.6195, -0.2643, -27.4378, 1.0, 4.5767, 0.2684, -28.6282, 1.0,
    4.4699, -3.3895, -30.6204, 1.0, 1.8125, -5.8448, -25.9695, 1.0, 1.9193, -2.1869, -23.9774, 1.0,
    5.8765, -1.6541, -25.1678, 1.0, -3.7263, 1.9956, -20.7352, 1.0, -5.1135, -5.5956, -28.2388, 1.0,
    -5.0067, -1.9377, -26.2467, 1.0, -1.0495, -1.4050, -27.4371, 1.0, -1.1563, -5.0629, -29.4292, 1.0,
    -3.8137, -7.5182, -24.7784, 1.0, 0.2503, -3.3276, -23.9766, 1.0, 0.1435, -6.9855, -25.9688, 1.0,
    -4.5209, -0.3826, -22.9609, 1.0, -4.4455, 2.1991, -21.5549, 1.0, -1.6526, 2.5750, -22.3950, 1.0,
    -1.7281, -0.0066, -23.8010, 1.0, -3.6036, -1.7395, -20.5186, 1.0, -3.5282, 0.8422, -19.1126, 1.0,
    -0.7353, 1.2181, -19.9528, 1.0, -0.8107, -1.3635, -21.3588, 1.0, -3.3029, 1.3693, -19.6080, 1.0,
    -2.0139, 1.5429, -19.9957, 1.0], f32).reshape(25, 4)
TEST_GLX_FACES = np.array([
    4, 5, 1, 5, 6, 1, 0, 1, 2, 13, 14, 11, 14, 12, 8, 8, 9, 10, 19, 20, 16, 20, 21, 16, 21, 22, 17, 22, 19, 18,
    15, 16, 17, 22, 21, 20, 0, 4, 1, 21, 17, 16, 13, 10, 9, 3, 0, 2, 8, 12, 9, 22, 18, 17, 10, 13, 11, 11, 14, 8,
    11, 8, 10, 15, 19, 16, 23, 24, 7, 6, 2, 1, 18, 15, 17, 19, 22, 20, 19, 15, 18], np.int32).reshape(27, 3)
TEST_GLX_MVP = np.array([
    -1.195982575416565, 1.350219488143921, 1.237614393234253, 30.956573486328125,
    -0.1888779103755951, -2.055802583694458, 2.06032657623291, 47.59274673461914,
    -1.0203083753585815, -0.42725738883018494, -0.519854724407196, 2.6755423545837402,
    -0.834797739982605, -0.3495742380619049, -0.42533570528030396, 7.643625259399414], f32).reshape(4, 4)
TEST_GLX_SIDE_MVP = np.array([
    -1.831691861152649, -1.1502554416656494, -0.3270684480667114, -11.764444351196289,
    1.391772985458374, -2.4397428035736084, 0.7858548760414124, 19.515047073364258,
    0.3260231614112854, -0.188545361161232, -1.1627495288848877, -21.932016372680664,
    0.2667462229728699, -0.1542643904685974, -0.9513405561447144, -12.489831924438477], f32).reshape(4, 4)
