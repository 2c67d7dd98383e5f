#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the REFERENCE'S OWN outputs.

The reference ships no tests; the only artefacts that pin its results are the renders in
/root/reference/demo/*.png (1920x1080, i.e. `-q`: 4000 spp, reference src/main.rs:633), written by
Camera::render as sqrt-gamma 8-bit (src/camera.rs:109-114).  Scenes 2, 4, 5, 6 have deterministic
geometry (scene 1's is random per run, scenes 3 and 7 have no surviving render).

Each fixture is the demo image linearised ((byte + 0.5) / 256)^2, box-downsampled 8x to 240x135 and stored
as float16, plus the fraction of clipped (byte == 255) source pixels per cell so tests can mask cells where
the 0.999 clip biases the mean.   Run here (needs /root/reference):  python tests/golden/make_golden.py
"""
import os
import numpy as np
from PIL import Image

SRC = "/root/reference/demo"
DST = os.path.dirname(os.path.abspath(__file__))
F = 8
for name, scene in [("earth", 2), ("lights", 4), ("bsdf", 5), ("scene6", 6)]:
    b = np.asarray(Image.open(os.path.join(SRC, name + ".png")).convert("RGB"), dtype=np.float64)
    lin = ((b + 0.5) / 256.0) ** 2
    h, w, _ = lin.shape
    cells = lin.reshape(h // F, F, w // F, F, 3).mean(axis=(1, 3))
    clipped = (b >= 255).any(axis=2).reshape(h // F, F, w // F, F).mean(axis=(1, 3))
    np.savez_compressed(os.path.join(DST, f"demo_scene{scene}.npz"), mean=cells.astype(np.float16), clipped=clipped.astype(np.float16),
                        width=np.int32(w), height=np.int32(h), factor=np.int32(F))
    print(name, cells.shape, "mean", cells.mean(), "clipped cells", (clipped > 0).mean())
