"""Shared helpers for the tests (fixtures loading, tolerances)."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_LARGE = os.path.join(ROOT, "tests", "golden_large")

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


def large_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_LARGE, "*.npz")))


def load_large(name):
    """Recipe-size fixture written by oracle/build_ref.py from the UNMODIFIED reference.  The weights are rebuilt from
    the seed (synth.make_model) and must reproduce the per-tensor checksums stored with the fixture."""
    from shallow_wavenet_b200 import synth
    z = np.load(os.path.join(GOLDEN_LARGE, name + ".npz"))
    kw = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg.")}
    kind = str(z["kind"])
    model = synth.make_model(kind, seed=int(z["weight_seed"]), **kw)
    sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    assert list(sd) == [str(k) for k in z["sd_keys"]], "state_dict layout changed since the fixture was written"
    sums = np.array([np.asarray(sd[k], np.float64).sum() for k in sd])
    assert np.array_equal(sums, z["sd_sums"]), "synth.make_model no longer reproduces the fixture's weights"
    cfg = SWNConfig(kind=kind, **kw)
    return dict(name=name, kind=kind, cfg=cfg, kw=kw, sd=sd, model=model, aux=z["aux"], noise=z["noise"], heads=z["heads"],
                samples=z["samples"], n_samples_list=[int(v) for v in z["n_samples_list"]])


def head_components(kind, kw):
    """Column ranges of one out_2 row: (loc, pre-sigmoid scale, LP coefficients) for CSWNV (cswnv_shift1.py:368-372),
    the 256 logits for DSWNV."""
    if kind == "dswnv":
        return {"logits": slice(0, kw.get("n_quantize", 256))}
    seg, lpc = kw.get("seg", 5), kw.get("lpc", 4)
    c = {"loc": slice(0, seg), "scale": slice(seg, 2 * seg)}
    if lpc > 0:
        c["lp"] = slice(2 * seg, 2 * seg + lpc)
    return c


def head_errors(a, b, kind, kw, floor=0.05):
    """Per-component scaled error max|a-b| / (max|b_component row| + floor)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    out = {}
    for name, sl in head_components(kind, kw).items():
        d = np.abs(a[..., sl] - b[..., sl])
        scale = np.abs(b[..., sl]).max(axis=-1, keepdims=True) + floor
        out[name] = float((d / scale).max()) if d.size else 0.0
    return out


GOLDEN_FORWARD = os.path.join(ROOT, "tests", "golden_forward")


def forward_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_FORWARD, "*.npz")))


def load_forward(name):
    """Fixture of the training-time forward written by tests/golden/make_forward_golden.py from the UNMODIFIED reference.
    Returns kind, kw, cfg, sd, aux, audio and the reference outputs (``ref``: name -> array)."""
    z = np.load(os.path.join(GOLDEN_FORWARD, name + ".npz"))
    if "base" in z.files:
        g = load_golden(str(z["base"]))
        kind, kw, cfg, sd, aux = g["kind"], g["kw"], g["cfg"], g["sd"], g["aux"]
    else:
        from shallow_wavenet_b200 import synth
        kw = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg.")}
        kind = str(z["kind"])
        model = synth.make_model(kind, seed=int(z["weight_seed"]), **kw)
        sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        sums = np.array([np.asarray(sd[k], np.float64).sum() for k in sd])
        assert np.array_equal(sums, z["sd_sums"]), "synth.make_model no longer reproduces the fixture's weights"
        cfg, aux = SWNConfig(kind=kind, **kw), z["aux"]
    return dict(name=name, kind=kind, kw=kw, cfg=cfg, sd=sd, aux=aux, audio=z["audio"],
                ref={k[4:]: z[k] for k in z.files if k.startswith("ref.")})


def scaled_err(a, b, floor=0.05):
    """max|a-b| / (max|b| over the last axis + floor)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.ndim == 2:
        a, b = a[..., None], b[..., None]
    return float((np.abs(a - b) / (np.abs(b).max(axis=-1, keepdims=True) + floor)).max())
