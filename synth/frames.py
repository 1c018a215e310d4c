"""Synthetic video frames and detections (DATA ONLY): inputs for the crop / stream tests, benches and goldens."""
import numpy as np


def synthetic_frame(seed=0, H=360, W=480):
    """deterministic test frame: smooth gradients + texture + noise (so that interpolation errors show)"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W]
    f = np.stack([128 + 100 * np.sin(x / 17.0) * np.cos(y / 23.0), (x * 255.0 / W + y) % 256, 255.0 * ((x // 8 + y // 8) % 2)], -1)
    f = f + rng.normal(0, 25, size=f.shape)
    return np.clip(f, 0, 255).astype(np.uint8)


def synthetic_boxes(seed=0, n=9, H=360, W=480):
    """detections [cx, cy, w, h] as float32-representable values, several partly outside the frame"""
    rng = np.random.default_rng(seed + 100)
    b = np.stack([rng.uniform(-20, W + 20, n), rng.uniform(-20, H + 20, n), rng.uniform(30, 400, n), rng.uniform(30, 400, n)], 1)
    b[0] = [W / 2, H / 2, 200, 200]
    if n > 1:
        b[1] = [10.5, 12.25, 150, 90]
    return b.astype(np.float32).astype(np.float64)
