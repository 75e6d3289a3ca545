# -*- coding: utf-8 -*-
"""Synthetic workloads of BASELINE.json (random-init weights, WORLD-style aux features).

There is no network for checkpoints or corpora, so the benchmark and the large parity tests
build their inputs here (SURVEY.md 8d "Synthetic inputs"): weights by ``initialize`` under
``torch.manual_seed(1)`` with a non-trivial standardiser and per-phase upsampling gains, aux
features laid out like feature_extract.py:300-309 ([uv, log-F0, coded-ap, mcep])."""
import numpy as np
import torch

from .nets import CSWNV, DSWNV, initialize

# ctor kwargs of the recipe models (egs/shallow-wavenet/run.sh:162-230), by sampling rate
RATE = {16000: dict(n_aux=53, upsampling_factor=80),
        22050: dict(n_aux=54, upsampling_factor=110),
        24000: dict(n_aux=55, upsampling_factor=120)}


def cswnv_kwargs(fs=22050, seg=5, lpc=4):
    return dict(hid_chn=192, skip_chn=256, aux_kernel_size=3, aux_dilation_size=2, dilation_depth=3,
                dilation_repeat=2, kernel_size=7, seg=seg, lpc=lpc, wav_conv_flag=True, aux_conv2d_flag=False,
                **RATE[fs])


def dswnv_kwargs(fs=16000):
    return dict(n_quantize=256, hid_chn=256, skip_chn=256, aux_kernel_size=3, aux_dilation_size=2,
                dilation_depth=3, dilation_repeat=2, kernel_size=7, wav_conv_flag=False, audio_in_flag=False,
                **RATE[fs])


def random_init(model, seed=1):
    """initialize() + frozen-standardiser-like scale_in + perturbed upsampling gains + small biases."""
    torch.manual_seed(seed)
    model.apply(initialize)
    g = torch.Generator().manual_seed(seed + 100)
    n_aux = model.n_aux
    with torch.no_grad():
        sigma = 0.5 + 1.5 * torch.rand(n_aux, generator=g)
        mean = torch.randn(n_aux, generator=g)
        model.scale_in.weight.copy_(torch.diag(1.0 / sigma).unsqueeze(2))
        model.scale_in.bias.copy_(-mean / sigma)
        w = model.upsampling.conv.weight
        w.copy_(1.0 + 0.1 * torch.randn(w.shape, generator=g))
        for n, p in model.named_parameters():
            if n.endswith(".bias") and not n.startswith(("scale_in", "upsampling")):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    return model.eval()


def make_model(kind, seed=1, **kw):
    model = (CSWNV if kind == "cswnv" else DSWNV)(**kw)
    return random_init(model, seed)


def synth_aux(n_aux, frames, B, seed=1234):
    """[B, n_aux, frames] float32: col0 uv in blocks, col1 smoothed log-F0, then coded ap, then mcep."""
    rng = np.random.default_rng(seed)
    a = np.zeros((B, n_aux, frames), np.float32)
    for b in range(B):
        t = 0
        uv = np.zeros(frames, np.float32)
        v = 1.0
        while t < frames:
            n = int(rng.integers(20, 61))
            uv[t:t + n] = v
            v = 1.0 - v
            t += n
        a[b, 0] = uv
        lf0 = rng.uniform(np.log(80.0), np.log(400.0), frames + 8)
        a[b, 1] = np.convolve(lf0, np.ones(9) / 9.0, mode="valid")[:frames]
        n_ap = 2 if n_aux >= 54 else 1
        a[b, 2:2 + n_ap] = rng.uniform(-20.0, 0.0, (n_ap, frames))
        n_mc = n_aux - 2 - n_ap
        mc = rng.standard_normal((n_mc, frames)) * (0.5 / (1.0 + np.arange(n_mc)))[:, None]
        mc[0] = rng.standard_normal(frames) - 4.0
        a[b, 2 + n_ap:] = mc
    return a
