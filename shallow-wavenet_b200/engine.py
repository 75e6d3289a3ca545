# -*- coding: utf-8 -*-
"""Host-side owner of one ``swn_handle``: pushes a ``state_dict`` through the C ABI and
drives conditioning + the persistent generation kernel.  torch is used for device memory and
the current stream only."""
import ctypes
import os

import numpy as np
import torch

from . import _abi


class Engine(object):
    def __init__(self, kind, cfg_kwargs, state_dict, device):
        lib = _abi.load()
        self.lib = lib
        self.kind = kind
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _abi.SwnError("the B200 generation path needs a CUDA device (call model.cuda()); "
                                "there is no CPU fallback")
        cfg = _abi.SwnConfig()
        cfg.kind = _abi.KIND_CSWNV if kind == "cswnv" else _abi.KIND_DSWNV
        for k, v in cfg_kwargs.items():
            setattr(cfg, k, int(v))
        h = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _abi.check(lib.swn_create(ctypes.byref(cfg), dev_index, ctypes.byref(h)))
        self.h = h
        self.dev_index = dev_index
        for key, val in state_dict.items():
            a = np.ascontiguousarray(val.detach().to("cpu", torch.float32).numpy())
            shape = (ctypes.c_int64 * a.ndim)(*a.shape)
            _abi.check(lib.swn_set_tensor(h, key.encode(), a.ctypes.data_as(ctypes.c_void_p), shape, a.ndim))
        _abi.check(lib.swn_finalize(h))
        self.info = self.get_info()

    def get_info(self):
        info = _abi.SwnInfo()
        _abi.check(self.lib.swn_get_info(self.h, ctypes.byref(info)))
        return info.as_dict()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def condition(self, aux):
        """aux: float32 CUDA tensor [B, n_aux, T_f]."""
        assert aux.is_cuda and aux.dtype == torch.float32 and aux.is_contiguous()
        B, _, Tf = aux.shape
        _abi.check(self.lib.swn_condition(self.h, ctypes.c_void_p(aux.data_ptr()), B, Tf, self._stream()))
        self._B = B

    def generate(self, n_samples, seed=0, noise=None, forced=None, seed_audio=None, return_heads=False,
                 batch_tile=0, sync=True, poll_mode=0, prof=None, prof_start=0, kernel_sel=0, n_samples_list=None,
                 helper_delay_ns=0):
        info = self.info
        seg, hd = info["eff_seg"], info["head_dim"]
        B = self._B
        n_gen = (int(n_samples) // seg) * seg
        is_c = self.kind == "cswnv"
        out = torch.empty((B, n_gen), dtype=torch.float32 if is_c else torch.int64, device=self.device)
        heads = torch.empty((B, n_gen // seg, hd), dtype=torch.float32, device=self.device) if return_heads else None
        a = _abi.SwnGenArgs()
        a.n_samples = int(n_samples)
        a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        keep = []

        def dev(t, dtype, shape):
            if t is None:
                return None
            t = torch.as_tensor(t).to(self.device, dtype).contiguous()
            assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
            keep.append(t)
            return t.data_ptr()
        a.noise_dev = dev(noise, torch.float32, (B, n_gen) if is_c else (B, n_gen, 256))
        a.forced_dev = dev(forced, torch.float32 if is_c else torch.int64, (B, n_gen))
        a.seed_dev = dev(seed_audio, torch.float32 if is_c else torch.int64, (B, seg))
        a.out_dev = out.data_ptr()
        a.head_dev = heads.data_ptr() if heads is not None else None
        a.batch_tile = int(batch_tile)
        a.debug_helper_delay_ns = int(helper_delay_ns)      # test hook (cluster kernels): late helper CTAs
        a.poll_mode = int(poll_mode)
        a.kernel_sel = int(kernel_sel) or int(os.environ.get("SWN_KERNEL_SEL", "0"))   # debugging aid
        if n_samples_list is not None:
            lens = np.ascontiguousarray(n_samples_list, np.int64)
            assert lens.shape == (B,), (lens.shape, B)
            keep.append(lens)
            a.n_samples_list = lens.ctypes.data_as(ctypes.c_void_p)
        if prof is not None:       # int64 CUDA tensor [count, grid, 32]
            keep.append(prof)
            a.prof_dev = prof.data_ptr()
            a.prof_start = int(prof_start)
            a.prof_count = int(prof.shape[0])
        _abi.check(self.lib.swn_generate(self.h, ctypes.byref(a), self._stream()))
        self._keep = keep
        if sync:
            self.sync()
        return out, heads

    def teacher_forced(self, samples, sync=True):
        """Per-step out_2 rows of the recurrence fed with `samples` (swn_teacher_forced): [B, n_gen // seg, head_dim]."""
        info = self.info
        seg, hd = info["eff_seg"], info["head_dim"]
        is_c = self.kind == "cswnv"
        t = torch.as_tensor(samples).to(self.device, torch.float32 if is_c else torch.int64).contiguous()
        assert t.dim() == 2 and t.shape[0] == self._B, (tuple(t.shape), self._B)
        n_gen = (int(t.shape[1]) // seg) * seg
        assert n_gen == t.shape[1], "the number of samples must be a multiple of seg"
        heads = torch.empty((self._B, n_gen // seg, hd), dtype=torch.float32, device=self.device)
        _abi.check(self.lib.swn_teacher_forced(self.h, t.data_ptr(), n_gen, heads.data_ptr(), self._stream()))
        self._keep = [t]
        if sync:
            self.sync()
        return heads

    def forward(self, audio, clip=False, sync=True):
        """swn_forward: the training-time / evaluation forward over whole utterances (time-parallel tcgen05 GEMMs).
        audio: CSWNV float [B, T_a]; DSWNV int64 [B, T_a] indices or float one-hot [B, Q, T_a].  Returns a dict of CUDA
        tensors: ``out`` [B, T, head_dim] raw out_2 rows and, for CSWNV, ``b`` / ``log_b`` (and the clipped pair)."""
        info = self.info
        seg, hd = info["eff_seg"], info["head_dim"]
        is_c = self.kind == "cswnv"
        B = self._B
        a = _abi.SwnFwdArgs()
        if is_c:
            t = torch.as_tensor(audio).to(self.device, torch.float32).contiguous()
            assert t.dim() == 2 and t.shape[0] == B, (tuple(t.shape), B)
            a.audio_onehot = 0
        elif torch.as_tensor(audio).dtype.is_floating_point:
            t = torch.as_tensor(audio).to(self.device, torch.float32).contiguous()
            assert t.dim() == 3 and t.shape[0] == B, (tuple(t.shape), B)
            a.audio_onehot = 1
        else:
            t = torch.as_tensor(audio).to(self.device, torch.int64).contiguous()
            assert t.dim() == 2 and t.shape[0] == B, (tuple(t.shape), B)
            a.audio_onehot = 0
        Ta = int(t.shape[-1])
        T = Ta - seg + 1
        res = {"out": torch.empty((B, T, hd), dtype=torch.float32, device=self.device)}
        a.audio_dev = t.data_ptr()
        a.T_a = Ta
        a.clip = 1 if clip else 0
        a.out_dev = res["out"].data_ptr()
        if is_c:
            for k in ("b", "log_b") + (("b_clip", "log_b_clip") if clip else ()):
                res[k] = torch.empty((B, T, seg), dtype=torch.float32, device=self.device)
                setattr(a, k + "_dev", res[k].data_ptr())
        _abi.check(self.lib.swn_forward(self.h, ctypes.byref(a), self._stream()))
        self._keep = [t]
        if sync:
            self.sync()
        return res

    def sync(self):
        _abi.check(self.lib.swn_sync(self.h, self._stream()))
        self.info = self.get_info()
        self._keep = None

    def generate_host(self, aux_host, n_samples, seed=0):
        """Host buffers in, host buffers out (the ABI's end-to-end entry point)."""
        aux_host = np.ascontiguousarray(aux_host, np.float32)
        B, _, Tf = aux_host.shape
        seg = self.info["eff_seg"]
        n_gen = (int(n_samples) // seg) * seg
        out = np.empty((B, n_gen), np.float32 if self.kind == "cswnv" else np.int64)
        _abi.check(self.lib.swn_generate_host(self.h, aux_host.ctypes.data_as(ctypes.c_void_p), B, Tf,
                                              int(n_samples), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                              out.ctypes.data_as(ctypes.c_void_p)))
        self.info = self.get_info()
        return out

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.swn_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def philox_uniform(seed, kind, b, n0, n, n_quantize=256):
    """Host replica of the uniforms the in-kernel sampler draws (tests)."""
    lib = _abi.load()
    k = _abi.KIND_CSWNV if kind == "cswnv" else _abi.KIND_DSWNV
    out = np.empty((n,) if kind == "cswnv" else (n, n_quantize), np.float32)
    _abi.check(lib.swn_philox_uniform(int(seed), k, n_quantize, int(b), int(n0), int(n),
                                      out.ctypes.data_as(ctypes.c_void_p)))
    return out


def measure_smem_bandwidth(device=0):
    """Aggregate shared-memory read bandwidth in GB/s (roofline denominator of the batch-1 step)."""
    lib = _abi.load()
    v = ctypes.c_double()
    _abi.check(lib.swn_measure_smem_bandwidth(int(device), ctypes.byref(v)))
    return v.value
