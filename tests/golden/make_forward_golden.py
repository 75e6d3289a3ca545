#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Fixtures of the training-time ``forward``, produced by the UNMODIFIED reference modules:
``CSWNV.forward(aux, audio)`` (src/nets/cswnv_shift1.py:191-267) and ``DSWNV.forward(audio, aux)``
(src/nets/dswnv.py:250-276) on the CPU.

Small cases reuse the weights and aux of the generation fixtures (tests/golden/<base>.npz); recipe-size cases rebuild
their weights from the seed like tests/golden_large (oracle/build_ref.py) and store per-tensor checksums.

Usage (build container only):  python tests/golden/make_forward_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden_forward")

from oracle import ref_runner as RR  # noqa: E402
from _util import load_golden  # noqa: E402

SMALL = ["cswnv_seg1_lpc0", "cswnv_seg5_lpc4_wav",   # (cswnv_seg2_lpc3 has 2 layers: the reference forward needs do_prob > 0 there, cswnv_shift1.py:222-225)
         "cswnv_seg5_lpc0_conv2d", "cswnv_seg5_lpc4_h64",
         "dswnv_base", "dswnv_wavconv_k6", "dswnv_audio_in"]
# name: (kind, fs, ctor overrides, B, frames, weight seed)
LARGE = {"fwd_large_cswnv_seg5": ("cswnv", 22050, dict(seg=5, lpc=4), 2, 4, 1),
         "fwd_large_dswnv": ("dswnv", 16000, dict(), 2, 5, 1)}


def run_reference(kind, kw, model, aux, audio):
    with torch.no_grad():
        if kind == "cswnv":
            r = model(torch.from_numpy(aux), torch.from_numpy(audio).unsqueeze(1))
            rc = model(torch.from_numpy(aux), torch.from_numpy(audio).unsqueeze(1), clip=True)
            names = ("mu", "b", "log_b", "a") if kw.get("lpc", 0) > 0 else ("mu", "b", "log_b")
            namesc = ("mu", "b", "b_clip", "log_b_clip", "a") if kw.get("lpc", 0) > 0 else ("mu", "b", "b_clip", "log_b_clip")
            out = {"ref." + n: t.numpy() for n, t in zip(names, r)}
            out.update({"ref." + n: t.numpy() for n, t in zip(namesc, rc) if n in ("b_clip", "log_b_clip")})
            return out
        onehot = torch.nn.functional.one_hot(torch.from_numpy(audio), kw["n_quantize"]).float().transpose(1, 2)
        return {"ref.logits": model(onehot, torch.from_numpy(aux)).numpy()}


def make(name, kind, kw, sd, aux, extra):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    U = kw["upsampling_factor"]
    seg = kw.get("seg", 1) if kind == "cswnv" else 1
    B, _, Tf = aux.shape
    Ta = Tf * U - seg
    if kind == "cswnv":
        audio = np.clip(0.3 * rng.standard_normal((B, Ta)), -1, 1).astype(np.float32)
    else:
        audio = rng.integers(0, kw["n_quantize"], (B, Ta)).astype(np.int64)
    model = RR.build_model(kind, kw, sd)
    fx = dict(kind=kind, audio=audio, **extra)
    fx.update(run_reference(kind, kw, model, aux, audio))
    os.makedirs(OUT, exist_ok=True)
    p = os.path.join(OUT, name + ".npz")
    np.savez_compressed(p, **fx)
    shapes = {k: v.shape for k, v in fx.items() if k.startswith("ref.")}
    print("%-28s B=%d T_a=%d %s %.0f KB" % (name, B, Ta, shapes, os.path.getsize(p) / 1024))


if __name__ == "__main__":
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    for base in SMALL:
        g = load_golden(base)
        make("fwd_" + base, g["kind"], g["kw"], g["sd"], g["aux"], dict(base=base))
    from shallow_wavenet_b200 import synth
    for name, (kind, fs, over, B, frames, seed) in LARGE.items():
        kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
        kw.update(over)
        ours = synth.make_model(kind, seed=seed, **kw)
        sd = {k: v.detach().numpy() for k, v in ours.state_dict().items()}
        aux = synth.synth_aux(kw["n_aux"], frames, B, seed=2468)
        extra = dict(fs=fs, weight_seed=seed, aux=aux, sd_keys=np.array(list(sd)),
                     sd_sums=np.array([np.asarray(v, np.float64).sum() for v in sd.values()], np.float64))
        extra.update({"cfg." + k: np.array(v) for k, v in kw.items()})
        make(name, kind, kw, sd, aux, extra)
