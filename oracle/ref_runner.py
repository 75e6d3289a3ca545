# -*- coding: utf-8 -*-
"""TEST / BENCHMARK INFRASTRUCTURE: runs the UNMODIFIED reference modules on the host CPU.

The reference is Python (``src/nets/{cswnv_shift1,dswnv}.py``), so the "reference build" is a
verbatim copy of those two files made by ``oracle/build_ref.py`` (called from
``__graft_entry__.build()`` in the build container) into the git-ignored ``oracle/_ref/``;
it travels to the GPU box with the snapshot like the built ``.so`` files do.  In the build
container the modules can also be imported straight from ``/root/reference/src/nets``.

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU legs (``cpu_baseline``,
``--impl reference``) may import this module; nothing under ``shallow-wavenet_b200/`` does.

Shims, all of them undone on exit:
  * ``torch.Tensor.cuda`` -> identity: ``cswnv_shift1.py:345/347`` hard-code ``.cuda()`` for the noise tensor and
    ``dswnv.py:89-90`` moves the one-hot to CUDA whenever a GPU is visible, while the model is timed on the CPU;
  * optional noise replay for parity fixtures: ``torch.Tensor.uniform_`` (Laplace) / ``torch.multinomial`` (softmax,
    whose ATen fast path is ``argmax(p / q)``, ``q ~ Exp(1)``).
"""
import contextlib
import importlib.util
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/src/nets"
_FILES = {"cswnv": "cswnv_shift1.py", "dswnv": "dswnv.py"}
_mods = {}


def ref_dir():
    """Directory holding the unmodified reference modules, or None (then callers fall back to the port)."""
    for d in (REF_COPY, REF_SRC):
        if all(os.path.isfile(os.path.join(d, f)) for f in _FILES.values()):
            return d
    return None


def available():
    return ref_dir() is not None


def module(kind):
    """The reference module (cswnv_shift1 / dswnv), loaded from its file without touching sys.path."""
    if kind not in _mods:
        d = ref_dir()
        if d is None:
            raise RuntimeError("reference modules not found: run `python __graft_entry__.py build` in the build container "
                               "(it copies them to oracle/_ref/)")
        name = "_swn_reference_" + kind
        spec = importlib.util.spec_from_file_location(name, os.path.join(d, _FILES[kind]))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        _mods[kind] = mod
    return _mods[kind]


@contextlib.contextmanager
def cpu_shim():
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def build_model(kind, kw, state_dict=None):
    """Reference ``CSWNV(**kw)`` / ``DSWNV(**kw)`` on the CPU in eval mode, optionally strictly loaded."""
    mod = module(kind)
    model = (mod.CSWNV if kind == "cswnv" else mod.DSWNV)(**kw)
    if state_dict is not None:
        sd = {k: (v.detach().cpu() if torch.is_tensor(v) else torch.from_numpy(np.asarray(v))) for k, v in state_dict.items()}
        model.load_state_dict(sd, strict=True)
    model.eval()
    for p in model.parameters():
        p.requires_grad = False
    return model


@contextlib.contextmanager
def replay_noise(kind, kw, noise):
    """Make the reference's sampler consume ``noise`` ([B, N] uniforms in (-0.4999, 0.5) / [B, N, Q] Exp(1) variates)."""
    cnt = [0]
    if kind == "cswnv":
        seg, lpc = kw.get("seg", 5), kw.get("lpc", 4)
        orig = torch.Tensor.uniform_

        def fake_uniform(self, a=0.0, b=1.0, **k2):
            i = cnt[0]
            w = 1 if lpc > 0 else seg                # one [B,1,1] draw per sample / one [B,1,seg] draw per step
            self.copy_(torch.from_numpy(np.ascontiguousarray(noise[:, i:i + w])).reshape(self.shape))
            cnt[0] += w
            return self
        torch.Tensor.uniform_ = fake_uniform
        try:
            yield cnt
        finally:
            torch.Tensor.uniform_ = orig
    else:
        orig = torch.multinomial

        def fake_multinomial(p, num_samples, replacement=False, **k2):
            assert num_samples == 1 and replacement
            q = torch.from_numpy(np.ascontiguousarray(noise[:, cnt[0]])).reshape(p.shape)
            cnt[0] += 1
            return torch.argmax(p / q, dim=-1, keepdim=True)
        torch.multinomial = fake_multinomial
        try:
            yield cnt
        finally:
            torch.multinomial = orig


def generate(kind, kw, model, aux, n_samples_list, noise=None, capture_heads=False, intervals=10 ** 9):
    """``model.batch_fast_generate`` exactly as the decode scripts call it (decode_cswnv_laplace-shift1.py:247-248,
    decode_dswnv_softmax.py:254-255).  Returns (list of arrays, heads [B, n_iter, head_dim] or None, seconds)."""
    B = aux.shape[0]
    if kind == "cswnv":
        x = torch.zeros(B, kw.get("seg", 5))
    else:
        x = torch.full((B, 1), kw.get("n_quantize", 256) // 2, dtype=torch.long)
    heads = []
    hook = None
    if capture_heads:
        hook = model.out_2.register_forward_hook(lambda m, i, o: heads.append(o[:, :, -1].detach().clone()))
    ctx = replay_noise(kind, kw, noise) if noise is not None else contextlib.nullcontext()
    try:
        with cpu_shim(), ctx, torch.no_grad():
            t0 = time.perf_counter()
            out = model.batch_fast_generate(x, torch.as_tensor(aux), list(n_samples_list), intervals)
            dt = time.perf_counter() - t0
    finally:
        if hook is not None:
            hook.remove()
    return out, (torch.stack(heads, 1).numpy() if capture_heads else None), dt


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"
