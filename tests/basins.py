"""Small deterministic basins shared by the tests (pure integer functions, no RNG state)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


from ocean_model_arch_b200.basins import frame_mask, island_mask  # noqa: E402,F401  (the generators live in the package)


def bs_mask():
    """The reference's data/BS/mask_bs4km.txt (289x163), packed by tests/golden/make_fixtures.py."""
    z = np.load(os.path.join(GOLDEN, "bs4km_mask.npz"))
    nx, ny = int(z["nx"]), int(z["ny"])
    return np.unpackbits(z["bits"])[: nx * ny].reshape(ny, nx).astype(np.int32)
