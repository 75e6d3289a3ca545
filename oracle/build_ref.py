#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""TEST / BENCHMARK INFRASTRUCTURE: the "build" of the reference for this path.

The reference is Python, so there is nothing to compile: ``make_ref_copy()`` copies the two
``src/nets`` modules VERBATIM from ``/root/reference`` into the git-ignored ``oracle/_ref/``
(never into history) with a manifest of their SHA-256, so that the GPU box - which has no
``/root/reference`` - can run the unmodified ``batch_fast_generate`` as the CPU arm of
``bench.py`` (``cpu_baseline.kind = "reference"``).

``make_large_fixtures()`` runs the unmodified reference at RECIPE size (H=192 seg 5 / H=256
DSWNV, k=7, L=6) on a few frames with replayed noise and writes ``tests/golden_large/*.npz``.
The weights are not stored (38 MB per model): they are ``synth.make_model(kind, seed)`` of the
product package, strictly loaded into the reference class, and the fixture keeps a float64
checksum per tensor so that the test can prove it rebuilt the same weights.

Usage (build container only):  python oracle/build_ref.py [copy] [fixtures] [--force]
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
LARGE = os.path.join(ROOT, "tests", "golden_large")

from oracle import ref_runner as RR  # noqa: E402

# name: (kind, fs, ctor overrides, frames per utterance, weight seed)
LARGE_CASES = {
    "large_cswnv_seg5_b2": ("cswnv", 22050, dict(seg=5, lpc=4), [3, 2], 1),
    "large_cswnv_seg1_b1": ("cswnv", 22050, dict(seg=1, lpc=4), [2], 1),
    "large_cswnv_seg5_b8": ("cswnv", 22050, dict(seg=5, lpc=4), [2, 2, 1, 2, 2, 2, 1, 2], 1),
    "large_dswnv_b1": ("dswnv", 16000, dict(), [2], 1),
    "large_dswnv_b2": ("dswnv", 16000, dict(), [2, 1], 1),
    "large_dswnv_b8": ("dswnv", 24000, dict(), [1] * 8, 1),
}


def make_ref_copy():
    if not os.path.isdir(RR.REF_SRC):
        return False
    os.makedirs(RR.REF_COPY, exist_ok=True)
    manifest = {}
    for f in RR._FILES.values():
        src, dst = os.path.join(RR.REF_SRC, f), os.path.join(RR.REF_COPY, f)
        shutil.copyfile(src, dst)
        manifest[f] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": RR.REF_SRC, "sha256": manifest, "note": "verbatim copies, git-ignored, never edited"},
              open(os.path.join(RR.REF_COPY, "MANIFEST.json"), "w"), indent=1)
    return True


def tensor_checksums(sd):
    return {k: float(np.asarray(v, np.float64).sum()) for k, v in sd.items()}


def make_case(name, kind, fs, over, frames, seed):
    import torch
    from shallow_wavenet_b200 import synth
    kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
    kw.update(over)
    ours = synth.make_model(kind, seed=seed, **kw)
    sd = {k: v.detach().numpy() for k, v in ours.state_dict().items()}
    ref = RR.build_model(kind, kw, sd)
    U = kw["upsampling_factor"]
    B, T = len(frames), max(frames)
    aux = synth.synth_aux(kw["n_aux"], T, B, seed=4321)
    for b, t in enumerate(frames):
        aux[b, :, t:] = 0.0                          # pad_list zero padding (decode_cswnv_laplace-shift1.py:30-48)
    n_list = [t * U for t in frames]
    N = max(n_list)
    rng = np.random.default_rng(seed + 7)
    seg = kw.get("seg", 1) if kind == "cswnv" else 1
    if kind == "cswnv":
        noise = rng.uniform(-0.4999, 0.5, size=(B, N)).astype(np.float32)
    else:
        noise = rng.exponential(1.0, size=(B, N, kw["n_quantize"])).astype(np.float32)
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    out, heads, dt = RR.generate(kind, kw, ref, aux, [N] * B, noise=noise, capture_heads=True)
    n_gen = (N // seg) * seg
    samples = np.stack([np.asarray(o)[:n_gen] for o in out])
    assert heads.shape[1] == N // seg
    fx = dict(kind=kind, name=name, fs=fs, weight_seed=seed, n_samples_list=np.array(n_list), aux=aux, noise=noise,
              heads=heads.astype(np.float32), samples=samples, receptive_field=ref.receptive_field,
              sd_keys=np.array(list(sd)), sd_sums=np.array([tensor_checksums(sd)[k] for k in sd], np.float64))
    for k, v in kw.items():
        fx["cfg." + k] = np.array(v)
    os.makedirs(LARGE, exist_ok=True)
    p = os.path.join(LARGE, name + ".npz")
    np.savez_compressed(p, **fx)
    print("%-22s B=%d N=%d seg=%d heads=%s  reference %.1f s  %.0f KB" % (name, B, N, seg, heads.shape, dt, os.path.getsize(p) / 1024))


def make_large_fixtures(force=False, only=()):
    if not RR.available():
        return False
    for name, (kind, fs, over, frames, seed) in LARGE_CASES.items():
        if only and name not in only:
            continue
        if force or not os.path.exists(os.path.join(LARGE, name + ".npz")):
            make_case(name, kind, fs, over, frames, seed)
    return True


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    force = "--force" in sys.argv
    what = [a for a in args if a in ("copy", "fixtures")] or ["copy", "fixtures"]
    only = [a for a in args if a in LARGE_CASES]
    if "copy" in what:
        print("oracle/_ref:", "copied" if make_ref_copy() else "no /root/reference here")
    if "fixtures" in what:
        make_large_fixtures(force, only)
