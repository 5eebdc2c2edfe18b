"""Deterministic synthetic basins (pure integer functions, no RNG state): the land masks of the BASELINE
configs that have no input file.  Used by bench.py and by the tests."""
import numpy as np


def frame_mask(nx, ny):
    m = np.ones((ny, nx), dtype=np.int32)
    m[2:ny - 2, 2:nx - 2] = 0
    return m


def island_mask(nx, ny, seed=20240229, ndisc=6, coast=True):
    """2-cell land frame + discs from an LCG + a sinusoidal coast."""
    m = frame_mask(nx, ny)
    s = seed
    for _ in range(ndisc):
        s = (1103515245 * s + 12345) % (1 << 31); cx = s % nx
        s = (1103515245 * s + 12345) % (1 << 31); cy = s % ny
        s = (1103515245 * s + 12345) % (1 << 31); r = 2 + s % max(3, min(nx, ny) // 10)
        x0, x1, y0, y1 = max(cx - r, 0), min(cx + r + 1, nx), max(cy - r, 0), min(cy + r + 1, ny)
        jj, ii = np.mgrid[y0:y1, x0:x1]                       # only the disc's bounding box
        m[y0:y1, x0:x1][(ii - cx) ** 2 + (jj - cy) ** 2 <= r * r] = 1
    if coast:
        depth = (2 + (ny // 8) * (1 + np.sin(np.arange(nx) * (6.0 / nx)))).astype(np.int64)
        m[np.arange(ny)[:, None] < depth[None, :]] = 1
    # never cover the centre of the Gaussian bump completely
    m[ny // 2 - 1:ny // 2 + 2, nx // 2 - 1:nx // 2 + 2] = 0
    m[:2, :] = 1; m[-2:, :] = 1; m[:, :2] = 1; m[:, -2:] = 1
    return m
