#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Fixtures of the decode DRIVERS, produced by running the reference's own scripts end to end.

``src/bin/decode_cswnv_laplace-shift1.py`` and ``src/bin/decode_dswnv_softmax.py`` are loaded
UNMODIFIED from /root/reference and their ``main()`` is executed on the CPU: argument parsing,
``torch.load(model.conf)`` (pickled argparse.Namespace), ``torch.load(checkpoint)["model"]``,
the n_gpus split of the list, ``decode_generator`` (sort by length, array_split, zero padding,
seed waveform), ``batch_fast_generate``, clipping / ``decode_mu_law`` and ``sf.write``.

What is stubbed (none of it is arithmetic of the path):
  * ``h5py``       - absent here; a File-like that serves datasets from ``<name>.h5`` files that
                     really are ``.npz`` archives (written by this script);
  * ``soundfile``  - absent; ``sf.write`` records (path, samples, fs, subtype);
  * ``torchvision.transforms.Compose`` / ``distutils.util.strtobool`` - absent, trivial;
  * ``mp.Process`` - runs its target inline (the closure target needs fork; same code path);
  * ``.cuda()`` on tensors and modules, ``torch.cuda.device`` - identity (no GPU here);
  * ``torch.load``  - ``weights_only=False`` (torch >= 2.6 refuses the pickled Namespace otherwise);
  * the sampler's noise - ``Tensor.uniform_`` / ``torch.multinomial`` replay per-utterance noise
    stored in the fixture (which utterances form the current batch is read off a wrapper around
    the script's own ``decode_generator``).

The checkpoint and model.conf are written with the layout of the training scripts
(train_cswnv_laplace-stftcmplx_shift1.py:162-182 ``save_checkpoint``, :293 ``torch.save(args, ...)``).

Usage (build container only):  python tests/golden/make_decode_golden.py
"""
import argparse
import contextlib
import importlib.util
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "golden_decode")

CASES = {
    "decode_cswnv": dict(
        script="decode_cswnv_laplace-shift1.py", kind="cswnv", fs=22050, batch_size=3, n_gpus=2, seed=31,
        conf=dict(n_aux=6, skip_chn=16, hid_chn=24, dilation_depth=2, dilation_repeat=2, kernel_size=3, aux_kernel_size=3,
                  aux_dilation_size=2, seg=5, lpc=4, aux_conv2d_flag=False, wav_conv_flag=True, upsampling_factor=10,
                  string_path="/feat_org_lf0"),
        frames=[4, 7, 3, 7, 5, 2, 6, 4, 5]),
    "decode_dswnv": dict(
        script="decode_dswnv_softmax.py", kind="dswnv", fs=16000, batch_size=4, n_gpus=1, seed=32,
        conf=dict(n_quantize=256, n_aux=5, skip_chn=16, hid_chn=16, dilation_depth=2, dilation_repeat=2, kernel_size=3,
                  aux_kernel_size=3, aux_dilation_size=2, audio_in=False, wav_conv_flag=False, upsampling_factor=8,
                  string_path="/feat_org_lf0"),
        frames=[3, 5, 2, 5, 4, 6]),
}


class _FakeH5File(object):
    def __init__(self, name, mode="r"):
        self.z = np.load(name)

    def __contains__(self, path):
        return path.lstrip("/") in self.z.files

    def __getitem__(self, path):
        return self.z[path.lstrip("/")]

    def close(self):
        self.z.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class _InlineProcess(object):
    def __init__(self, target=None, args=(), kwargs=None):
        self.target, self.args, self.kwargs = target, args, kwargs or {}

    def start(self):
        self.target(*self.args, **self.kwargs)

    def join(self, timeout=None):
        return None


def _install_stubs(written):
    h5 = types.ModuleType("h5py")
    h5.File = _FakeH5File
    sf = types.ModuleType("soundfile")
    sf.write = lambda path, data, fs, subtype=None: written.append((path, np.array(data), int(fs), subtype))
    tv = types.ModuleType("torchvision")
    tr = types.ModuleType("torchvision.transforms")

    class Compose(object):
        def __init__(self, fs):
            self.fs = fs

        def __call__(self, x):
            for f in self.fs:
                x = f(x)
            return x
    tr.Compose = Compose
    tv.transforms = tr
    du = types.ModuleType("distutils")
    duu = types.ModuleType("distutils.util")
    duu.strtobool = lambda s: int(str(s).lower() in ("y", "yes", "t", "true", "on", "1"))
    du.util = duu
    for name, mod in (("h5py", h5), ("soundfile", sf), ("torchvision", tv), ("torchvision.transforms", tr),
                      ("distutils", du), ("distutils.util", duu)):
        sys.modules[name] = mod


def build_reference_model(kind, conf, seed):
    sys.path.insert(0, os.path.join(REF, "nets"))
    torch.manual_seed(seed)
    if kind == "cswnv":
        import cswnv_shift1 as ref
        kw = {k: v for k, v in conf.items() if k != "string_path"}
        model = ref.CSWNV(**kw)
    else:
        import dswnv as ref
        kw = {k: v for k, v in conf.items() if k not in ("string_path", "audio_in")}
        kw["audio_in_flag"] = conf["audio_in"]
        model = ref.DSWNV(**kw)
    model.apply(ref.initialize)
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        n_aux = conf["n_aux"]
        sigma = 0.5 + 1.5 * torch.rand(n_aux, generator=g)
        mean = torch.randn(n_aux, generator=g)
        model.scale_in.weight.copy_(torch.diag(1.0 / sigma).unsqueeze(2))
        model.scale_in.bias.copy_(-mean / sigma)
        model.upsampling.conv.weight.add_(0.1 * torch.randn(model.upsampling.conv.weight.shape, generator=g))
        for n, p in model.named_parameters():
            if n.endswith(".bias") and not n.startswith(("scale_in", "upsampling")):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return model, kw


def run_case(name, c):
    kind, conf = c["kind"], c["conf"]
    rng = np.random.default_rng(c["seed"])
    work = tempfile.mkdtemp(prefix="swn_decode_golden_")
    try:
        model, ctor_kw = build_reference_model(kind, conf, c["seed"])
        # ---- files in the reference's formats
        ns = argparse.Namespace(**conf)
        torch.save(ns, os.path.join(work, "model.conf"))                                   # train_cswnv...py:293
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        ckpt = {"model": model.state_dict(), "optimizer": opt.state_dict(), "numpy_random_state": np.random.get_state(),
                "torch_random_state": torch.get_rng_state(), "iterations": 7}               # :171-176
        torch.save(ckpt, os.path.join(work, "checkpoint-7.pkl"))
        featdir = os.path.join(work, "feats")
        os.makedirs(featdir)
        U = conf["upsampling_factor"]
        n_max = max(c["frames"]) * U
        feats, noise = {}, {}
        for i, t in enumerate(c["frames"]):
            fid = "utt%02d" % i
            a = rng.standard_normal((t, conf["n_aux"])).astype(np.float32)
            a[:, 0] = (rng.random(t) > 0.4)
            a[:, 1] = np.log(rng.uniform(80, 400, t))
            feats[fid] = a
            with open(os.path.join(featdir, fid + ".h5"), "wb") as f:                       # an .npz behind an .h5 name
                np.savez(f, **{conf["string_path"].lstrip("/"): a})
            if kind == "cswnv":
                noise[fid] = rng.uniform(-0.4999, 0.5, n_max).astype(np.float32)
            else:
                noise[fid] = rng.exponential(1.0, (n_max, conf["n_quantize"])).astype(np.float32)
        # ---- the reference script, unmodified, with the stubs of the module docstring
        written = []
        _install_stubs(written)
        for sub in ("utils", "nets", "bin"):
            sys.path.insert(0, os.path.join(REF, sub))
        spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, "bin", c["script"]))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.mp.Process = _InlineProcess
        current = {"ids": None, "pos": 0}
        batches = []
        real_gen = mod.decode_generator

        def gen_wrapper(*a, **k):
            for feat_ids, payload in real_gen(*a, **k):
                current["ids"], current["pos"] = list(feat_ids), 0
                batches.append(list(feat_ids))
                yield feat_ids, payload
            current["ids"] = None            # between lists the model of the next process is constructed (its init draws)
        mod.decode_generator = gen_wrapper
        seg, lpc = conf.get("seg", 1), conf.get("lpc", 0)

        real_uniform = torch.Tensor.uniform_

        def fake_uniform(self, a=0.0, b=1.0, **k2):
            if current["ids"] is None:       # parameter initialisation of the model the script constructs
                return real_uniform(self, a, b, **k2)
            w = 1 if lpc > 0 else seg
            i = current["pos"]
            rows = np.stack([noise[f][i:i + w] for f in current["ids"]])
            self.copy_(torch.from_numpy(rows).reshape(self.shape))
            current["pos"] += w
            return self

        def fake_multinomial(p, num_samples, replacement=False, **k2):
            q = torch.from_numpy(np.stack([noise[f][current["pos"]] for f in current["ids"]]))
            current["pos"] += 1
            return torch.argmax(p / q, dim=-1, keepdim=True)
        orig = (torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.device, torch.load, torch.Tensor.uniform_, torch.multinomial,
                sys.argv)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.device = lambda dev: contextlib.nullcontext()
        real_load = torch.load
        torch.load = lambda f, *a, **k: real_load(f, *a, **dict(k, weights_only=False))
        torch.Tensor.uniform_ = fake_uniform
        torch.multinomial = fake_multinomial
        outdir = os.path.join(work, "out")
        sys.argv = [c["script"], "--feats", featdir, "--checkpoint", os.path.join(work, "checkpoint-7.pkl"), "--config",
                    os.path.join(work, "model.conf"), "--outdir", outdir, "--fs", str(c["fs"]), "--batch_size",
                    str(c["batch_size"]), "--n_gpus", str(c["n_gpus"]), "--verbose", "0"]
        try:
            mod.main()
        finally:
            (torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.device, torch.load, torch.Tensor.uniform_, torch.multinomial,
             sys.argv) = orig
        assert len(written) == len(c["frames"]), (len(written), len(c["frames"]))
        fx = dict(kind=kind, fs=c["fs"], batch_size=c["batch_size"], n_gpus=c["n_gpus"],
                  order=np.array([os.path.basename(p)[:-4] for p, _, _, _ in written]),
                  batches=np.array(["|".join(b) for b in batches]))
        for k, v in conf.items():
            fx["conf." + k] = np.array(v)
        for k, v in model.state_dict().items():
            fx["sd." + k] = v.numpy()
        for fid in feats:
            fx["feat." + fid] = feats[fid]
            fx["noise." + fid] = noise[fid]
        for p, data, fs, subtype in written:
            assert fs == c["fs"] and subtype == "PCM_16"
            fx["wav." + os.path.basename(p)[:-4]] = data
        os.makedirs(OUT, exist_ok=True)
        dst = os.path.join(OUT, name + ".npz")
        np.savez_compressed(dst, **fx)
        print("%-14s %d utterances, batches %s, %.0f KB" % (name, len(written), [len(b) for b in batches],
                                                              os.path.getsize(dst) / 1024))
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    for name, c in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        run_case(name, c)
