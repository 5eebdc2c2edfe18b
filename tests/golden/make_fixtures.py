"""Regenerates the committed fixtures under tests/golden/.  Run in the build container only
(it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_fixtures.py

 - bs4km_mask.npz : the reference's shipped Black Sea mask data/BS/mask_bs4km.txt (289 x 163,
                    tools/io.f90:61-71 orientation), bit-packed;
 - oracle_golden.json : sha256 + sum/min/max of ssh / ubrtr / vbrtr of the CPU oracle after N steps
                    on small basins (bitwise pins; the arrays themselves are reproduced live).  The
                    reference ships NO golden vectors for this path (parity unpinned); these pin
                    the oracle against regressions and travel to the GPU box for the -m gpu tests.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

REF = "/root/reference"


def pack_bs_mask():
    nx, ny = 289, 163
    with open(os.path.join(REF, "data/BS/mask_bs4km.txt")) as f:
        f.readline()
        rows = [f.readline().rstrip("\n") for _ in range(ny)]
    m = np.empty((ny, nx), dtype=np.uint8)
    for i, row in enumerate(rows):
        m[ny - 1 - i, :] = np.frombuffer(row[:nx].encode(), dtype=np.uint8) - ord("0")
    assert set(np.unique(m)) <= {0, 1}
    np.savez_compressed(os.path.join(HERE, "bs4km_mask.npz"), nx=nx, ny=ny, bits=np.packbits(m.ravel()))
    print("bs4km sea cells:", int((m == 0).sum()))


CASES = {
    "rect32x24": dict(nx=36, ny=28, mask=None, steps=(1, 10, 100)),
    "islands64x48": dict(nx=68, ny=52, mask="islands", steps=(1, 10, 100)),
    "bs4km": dict(nx=289, ny=163, mask="bs", steps=(1, 10, 100),
                  cfg=dict(dxst=0.05, dyst=0.04, rlon=27.525, rlat=40.94)),
    "rect_cart_notrans": dict(nx=40, ny=30, mask=None, steps=(50,),
                              cfg=dict(curve_grid=0, trans_terms=0, dxst=0.01, dyst=0.01)),
    "islands_mu_rdiss": dict(nx=68, ny=52, mask="islands", steps=(100,),
                             cfg=dict(keep_mu=1, r_diss=5e-6, lvisc_2=1.0e3)),
}


def digest(a):
    import hashlib
    a = np.ascontiguousarray(a)
    return dict(sha256=hashlib.sha256(a.tobytes()).hexdigest(), sum=float(a.sum()), min=float(a.min()),
                max=float(a.max()))


def case_mask(c):
    import basins
    if c["mask"] == "islands":
        return basins.island_mask(c["nx"], c["ny"])
    if c["mask"] == "bs":
        return basins.bs_mask()
    return None


def oracle_goldens():
    import json
    from oracle_lib import OracleModel, make_config
    out = {}
    for name, c in CASES.items():
        m = OracleModel(make_config(c["nx"], c["ny"], **c.get("cfg", {})), case_mask(c))
        done = 0
        for s in c["steps"]:
            m.step(s - done)
            done = s
            for f in ("ssh", "ubrtr", "vbrtr"):
                out[f"{name}/{f}/{s}"] = digest(m.get(f))
        print(name, "ssh max after", done, "steps:", out[f"{name}/ssh/{done}"]["max"])
    with open(os.path.join(HERE, "oracle_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    pack_bs_mask()
    oracle_goldens()
