"""Shared helpers for the tests (fixtures loading, tolerances)."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle.swn_oracle import SWNConfig  # noqa: E402

# fp32 parity tolerance of the north-star ("e.g. 1e-4 relative"): a head value may
# differ by RTOL relative to the largest magnitude of its row plus ATOL absolute.
RTOL = 1e-4
ATOL = 1e-5


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg.")}
    sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
    kind = str(z["kind"])
    cfg = SWNConfig(kind=kind, **kw)
    return dict(name=name, kind=kind, cfg=cfg, kw=kw, sd=sd, aux=z["aux"], noise=z["noise"], heads=z["heads"],
                samples=z["samples"], n_samples_list=[int(v) for v in z["n_samples_list"]])


def assert_close(a, b, what="", rtol=RTOL, atol=ATOL):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-30)
    err = np.abs(a - b) / (atol / rtol + scale)
    worst = float(err.max()) if err.size else 0.0
    assert worst <= rtol, "%s: max scaled err %.3e > %.1e (abs %.3e)" % (what, worst, rtol, np.abs(a - b).max())
    return worst
