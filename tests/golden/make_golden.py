#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Generate golden fixtures by running the UNMODIFIED reference modules.

Runs only in the build container (needs /root/reference, read-only).  The GPU
box never runs this; it reads the committed ``*.npz`` files next to it.

For every case below the reference ``CSWNV`` / ``DSWNV`` is constructed with
``model.apply(initialize)`` under a fixed seed (as train_cswnv...py:343 does),
its ``scale_in`` / ``upsampling`` weights are perturbed so that every term of
the conditioning path is exercised, and ``batch_fast_generate`` is run on CPU
with
  * ``torch.Tensor.cuda`` patched to identity (cswnv_shift1.py:345/347 hard-code .cuda()),
  * ``torch.Tensor.uniform_`` / ``torch.multinomial`` patched to replay recorded noise,
  * a forward hook on ``out_2`` capturing the head output of every loop step.
Stored per case: config, state_dict, aux, n_samples_list, noise, per-step head
outputs, generated samples (and the teacher-forced ``forward`` is NOT used: the
generation path has its own warm-up semantics).

Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src/nets"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (kind, ctor kwargs, B, frames per utterance, seed)
    "cswnv_seg1_lpc0": ("cswnv", dict(n_aux=5, hid_chn=16, skip_chn=24, aux_kernel_size=3, aux_dilation_size=2,
                                      dilation_depth=2, dilation_repeat=2, kernel_size=3, upsampling_factor=6,
                                      seg=1, lpc=0, wav_conv_flag=False), [5, 4], 11),
    "cswnv_seg5_lpc4_wav": ("cswnv", dict(n_aux=6, hid_chn=24, skip_chn=16, aux_kernel_size=3, aux_dilation_size=2,
                                          dilation_depth=3, dilation_repeat=2, kernel_size=3, upsampling_factor=10,
                                          seg=5, lpc=4, wav_conv_flag=True), [6, 4, 5], 12),
    "cswnv_seg2_lpc3": ("cswnv", dict(n_aux=4, hid_chn=16, skip_chn=16, aux_kernel_size=3, aux_dilation_size=1,
                                      dilation_depth=2, dilation_repeat=1, kernel_size=4, upsampling_factor=8,
                                      seg=2, lpc=3, wav_conv_flag=True), [5], 13),
    "cswnv_seg5_lpc0_conv2d": ("cswnv", dict(n_aux=4, hid_chn=16, skip_chn=16, aux_kernel_size=3, aux_dilation_size=2,
                                             dilation_depth=2, dilation_repeat=2, kernel_size=3, upsampling_factor=10,
                                             seg=5, lpc=0, wav_conv_flag=True, aux_conv2d_flag=True), [4, 4], 14),
    "cswnv_seg5_lpc4_h64": ("cswnv", dict(n_aux=6, hid_chn=64, skip_chn=32, aux_kernel_size=3, aux_dilation_size=2,
                                          dilation_depth=3, dilation_repeat=2, kernel_size=7, upsampling_factor=20,
                                          seg=5, lpc=4, wav_conv_flag=True), [3, 2], 15),
    "dswnv_base": ("dswnv", dict(n_quantize=256, n_aux=5, hid_chn=16, skip_chn=16, aux_kernel_size=3,
                                 aux_dilation_size=2, dilation_depth=2, dilation_repeat=2, kernel_size=3,
                                 upsampling_factor=8), [5, 3], 21),
    "dswnv_wavconv_k6": ("dswnv", dict(n_quantize=256, n_aux=4, hid_chn=24, skip_chn=16, aux_kernel_size=3,
                                       aux_dilation_size=2, dilation_depth=2, dilation_repeat=3, kernel_size=6,
                                       upsampling_factor=10, wav_conv_flag=True), [4], 22),
    "dswnv_audio_in": ("dswnv", dict(n_quantize=256, n_aux=4, hid_chn=16, skip_chn=16, aux_kernel_size=3,
                                     aux_dilation_size=2, dilation_depth=2, dilation_repeat=1, kernel_size=3,
                                     upsampling_factor=8, audio_in_flag=True), [4, 4], 23),
}


def synth_aux(rng, n_aux, frames):
    """WORLD-style layout [uv, log-F0, coded ap..., mcep...] (feature_extract.py:300-309)."""
    T = max(frames)
    out = np.zeros((len(frames), n_aux, T), np.float32)
    for b, t in enumerate(frames):
        a = rng.standard_normal((n_aux, t)).astype(np.float32)
        a[0] = (rng.random(t) > 0.4).astype(np.float32)
        a[1] = np.log(rng.uniform(80, 400, t)).astype(np.float32)
        if n_aux > 2:
            a[2] = rng.uniform(-20, 0, t)
        out[b, :, :t] = a            # zero padded like pad_list (decode_cswnv...py:30-48)
    return out


def build(kind, kw, seed):
    torch.manual_seed(seed)
    if kind == "cswnv":
        import cswnv_shift1 as ref
        model = ref.CSWNV(**kw)
    else:
        import dswnv as ref
        model = ref.DSWNV(**kw)
    model.apply(ref.initialize)
    g = torch.Generator().manual_seed(seed + 100)
    n_aux = kw["n_aux"]
    with torch.no_grad():
        sigma = 0.5 + 1.5 * torch.rand(n_aux, generator=g)
        mean = torch.randn(n_aux, generator=g)
        model.scale_in.weight.copy_(torch.diag(1.0 / sigma).unsqueeze(2))
        model.scale_in.bias.copy_(-mean / sigma)
        model.upsampling.conv.weight.add_(0.1 * torch.randn(model.upsampling.conv.weight.shape, generator=g))
        model.upsampling.conv.bias.add_(0.05)
        for n, p in model.named_parameters():    # biases are zero after initialize(): make them count
            if n.endswith(".bias") and not n.startswith(("scale_in", "upsampling")):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    model.eval()
    return model


def run_case(name, kind, kw, frames, seed):
    model = build(kind, kw, seed)
    rng = np.random.default_rng(seed)
    aux = synth_aux(rng, kw["n_aux"], frames)
    U = kw["upsampling_factor"]
    n_list = [f * U for f in frames]
    B, N = len(frames), max(n_list)
    seg = kw.get("seg", 1) if kind == "cswnv" else 1

    heads = []
    hook = model.out_2.register_forward_hook(lambda m, i, o: heads.append(o[:, :, -1].detach().clone()))
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    if kind == "cswnv":
        noise = rng.uniform(-0.4999, 0.5, size=(B, N)).astype(np.float32)
        lpc = kw.get("lpc", 0)
        cnt = [0]
        orig = torch.Tensor.uniform_

        def fake_uniform(self, a=0.0, b=1.0, **k2):
            i = cnt[0]
            if lpc > 0:                      # one [B,1,1] draw per sample
                self.copy_(torch.from_numpy(noise[:, i:i + 1]).reshape(self.shape))
                cnt[0] += 1
            else:                            # one [B,1,seg] draw per step
                self.copy_(torch.from_numpy(noise[:, i:i + seg]).reshape(self.shape))
                cnt[0] += seg
            return self
        torch.Tensor.uniform_ = fake_uniform
        try:
            out = model.batch_fast_generate(torch.zeros(B, seg), torch.from_numpy(aux), [N] * B, 10 ** 9)
        finally:
            torch.Tensor.uniform_ = orig
        samples = np.zeros((B, (N // seg) * seg), np.float32)
    else:
        Q = kw["n_quantize"]
        noise = rng.exponential(1.0, size=(B, N, Q)).astype(np.float32)
        cnt = [0]
        orig = torch.multinomial

        # OneHotCategorical.sample -> torch.multinomial(p, 1, True), whose ATen fast path is
        # argmax(p / q), q ~ Exp(1) drawn inside C++ (checked here: identical under the same
        # seed, 20/20 trials).  Replay recorded q instead of the generator's.
        def fake_multinomial(p, num_samples, replacement=False, **k2):
            assert num_samples == 1 and replacement
            q = torch.from_numpy(noise[:, cnt[0]]).reshape(p.shape)
            cnt[0] += 1
            return torch.argmax(p / q, dim=-1, keepdim=True)
        torch.multinomial = fake_multinomial
        try:
            out = model.batch_fast_generate(torch.full((B, 1), Q // 2, dtype=torch.long),
                                            torch.from_numpy(aux), [N] * B, 10 ** 9)
        finally:
            torch.multinomial = orig
        assert cnt[0] == N
        samples = np.zeros((B, N), np.int64)
    torch.Tensor.cuda = orig_cuda
    hook.remove()
    # the loop runs max length for every utterance (cswnv_shift1.py:291-295); it is
    # called with [N]*B so the full rows come back; the per-utterance truncation
    # (cswnv_shift1.py:426-428) is exercised separately from n_samples_list.
    heads = torch.stack(heads, 1).numpy()  # [B, n_iter, head_dim]
    assert heads.shape[1] == N // seg, (heads.shape, N, seg)
    lens = np.array([len(o) for o in out])
    for b, o in enumerate(out):
        samples[b, :len(o)] = o
    fx = dict(kind=kind, name=name, n_samples_list=np.array(n_list), aux=aux, noise=noise, heads=heads,
              samples=samples, sample_lens=lens)
    for k, v in kw.items():
        fx["cfg." + k] = np.array(v)
    for k, v in model.state_dict().items():
        fx["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
    print("%-26s B=%d N=%d seg=%d heads=%s R=%d  %.0f KB" % (
        name, B, N, seg, heads.shape, model.receptive_field,
        os.path.getsize(os.path.join(HERE, name + ".npz")) / 1024))


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, (kind, kw, frames, seed) in CASES.items():
        if only and name not in only:
            continue
        run_case(name, kind, kw, frames, seed)
