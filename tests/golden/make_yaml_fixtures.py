"""Extracts small fixtures from the reference's own scene files (tracks/*.yaml: BASELINE configs 1-3).

    python tests/golden/make_yaml_fixtures.py          # needs /root/reference (only present in the build container)

The YAMLs are the reference's real camera tracks (4x4 float32 projections written by io_export_tracks.py) and
bundle points; the clips they belong to are missing upstream (.MISSING_LARGE_BLOBS), so the tests synthesise
stand-in frames through these cameras.  Only a handful of cameras per scene is kept (the .npz stay tiny)."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/tracks"

for name, picks in [("zatisi", [0, 1, 2, 3, 60, 61, 62]), ("koberec", [0, 1, 2, 3, 80, 81, 82]), ("koule-tr", [0, 1, 2, 3, 15, 16, 17])]:
    fs = cv2.FileStorage(os.path.join(REF, name + ".yaml"), cv2.FILE_STORAGE_READ)
    clip = fs.getNode("clip")
    W, H = int(clip.getNode("width").real()), int(clip.getNode("height").real())
    cams = fs.getNode("camera")
    P, near, far, frame_no = [], [], [], []
    for i in picks:
        c = cams.at(i)
        P.append(c.getNode("projection").mat().astype(np.float32))
        near.append(c.getNode("near").real())
        far.append(c.getNode("far").real())
        frame_no.append(int(c.getNode("frame").real()))
    tr = fs.getNode("tracks")
    bundles = np.stack([tr.at(i).getNode("bundle").mat().ravel() for i in range(tr.size())]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, f"tracks_{name}.npz"), W=W, H=H, cameras=np.stack(P), near=np.array(near), far=np.array(far),
                        frames=np.array(frame_no), bundles=bundles, n_cameras_total=cams.size())
    print(name, W, H, "cameras", cams.size(), "kept", len(picks), "bundles", bundles.shape)
