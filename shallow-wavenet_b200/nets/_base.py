# -*- coding: utf-8 -*-
"""Shared host-side plumbing of the two vocoder classes.

The classes in ``cswnv_shift1.py`` / ``dswnv.py`` mirror the reference's ``src/nets`` modules
at the boundary only: same constructor keywords, same public attributes, same parameter tree
(``state_dict`` keys and shapes) and the same ``batch_fast_generate``
call.  The parameters are plain ``nn.Parameter`` holders; all arithmetic of the generation
path runs in the CUDA library behind ``include/swn_b200.h``.
"""
import logging
import time

import numpy as np
import torch
from torch import nn

from ..engine import Engine


class _ParamConv(nn.Module):
    """Holder with ``weight`` / ``bias`` of a given shape (stands where the reference has an
    nn.Conv1d / nn.Conv2d / nn.ConvTranspose2d)."""

    def __init__(self, weight_shape, transposed=False):
        super(_ParamConv, self).__init__()
        self.transposed = transposed
        self.weight = nn.Parameter(torch.zeros(*weight_shape))
        self.bias = nn.Parameter(torch.zeros(weight_shape[1] if transposed else weight_shape[0]))

    def extra_repr(self):
        return "weight=%s" % (tuple(self.weight.shape),)


class _Wrapped(nn.Module):
    """``<name>.conv`` level of the reference's UpSampling / CausalConv1d wrappers."""

    def __init__(self, conv, **attrs):
        super(_Wrapped, self).__init__()
        self.conv = conv
        for k, v in attrs.items():
            setattr(self, k, v)


def initialize(m):
    """Xavier-uniform for the conv holders, ones for the upsampling deconvolution, zero
    biases -- what ``model.apply(initialize)`` does in the reference
    (src/nets/cswnv_shift1.py:20-34, src/nets/dswnv.py:50-64)."""
    if isinstance(m, _ParamConv):
        with torch.no_grad():
            if m.transposed:
                m.weight.fill_(1.0)
            else:
                nn.init.xavier_uniform_(m.weight)
            m.bias.zero_()


class _SWNBase(nn.Module):
    _kind = None

    # ------------------------------------------------------------------ construction
    def _build_aux(self):
        """Conditioning parameters (cswnv_shift1.py:151-157, dswnv.py:210-217)."""
        ak = self.aux_kernel_size
        self.scale_in = _ParamConv((self.n_aux, self.n_aux, 1))
        layers = nn.ModuleList()
        for i in range(self.aux_dilation_size):
            layers.append(_ParamConv((self.n_aux * ak ** (i + 1), self.n_aux * ak ** i, ak)))
        self.conv_aux = nn.Module()
        self.conv_aux.conv = layers
        self.conv_aux.rec_field = ak ** self.aux_dilation_size
        self.upsampling = _Wrapped(_ParamConv((1, 1, 1, self.upsampling_factor), transposed=True),
                                   upsampling_factor=self.upsampling_factor)
        if self.do_prob > 0:
            self.aux_drop = nn.Dropout(p=self.do_prob)

    def _build_audio_in(self, in_audio_ch):
        """wav_conv / causal input conv (cswnv_shift1.py:163-167, dswnv.py:222-226)."""
        k, H = self.kernel_size, self.n_hidch
        if self.wav_conv_flag:
            self.wav_conv = _ParamConv((H, in_audio_ch, 1))
            self.causal = _Wrapped(_ParamConv((H, H, k)), dil_fact=0, dilation=1, padding=k - 1)
        else:
            self.causal = _Wrapped(_ParamConv((H, in_audio_ch, k)), dil_fact=0, dilation=1, padding=k - 1)

    def _build_cells(self, in_x_dim):
        k, H, S = self.kernel_size, self.n_hidch, self.n_skipch
        self.padding = []
        self.dil_facts = [i for i in range(self.dilation_depth)] * self.dilation_repeat
        self.in_x = nn.ModuleList()
        self.dil_h = nn.ModuleList()
        self.out_skip = nn.ModuleList()
        for d in self.dil_facts:
            self.in_x.append(_ParamConv((2 * H, in_x_dim, 1)))
            pad = k ** (d + 1) - k ** d
            self.dil_h.append(_Wrapped(_ParamConv((2 * H, H, k)), dil_fact=d, dilation=k ** d, padding=pad))
            self.padding.append(pad)
            self.out_skip.append(_ParamConv((S, H, 1)))
        self.receptive_field = sum(self.padding) + k - 1
        if self.do_prob > 0:
            self.dcrnn_drop = nn.Dropout(p=self.do_prob)

    # ------------------------------------------------------------------ engine management
    def _cfg_kwargs(self):
        raise NotImplementedError

    #: also key the packed-weight cache on the CONTENT of the parameters (one float64 sum per tensor, i.e. ~50 small torch
    #: reductions on the device per generate call).  Off by default: the generation path launches no torch kernel.
    check_weight_content = False

    def _weights_version(self):
        """Key of the packed-weight cache: ``p._version`` and the storage pointers of every parameter.  Edits made through
        ``p.data`` in place (weight-norm removal, EMA swaps: ``.data`` is a detached alias with its own version counter)
        change neither: call ``refresh_engine()`` after such an edit, or set ``check_weight_content = True`` to have the
        key carry a content checksum as well."""
        params = list(self.parameters())
        key = (tuple(p._version for p in params), tuple(p.data_ptr() for p in params), str(params[0].device))
        if self.check_weight_content:
            with torch.no_grad():
                key += (torch.stack([p.detach().sum(dtype=torch.float64) for p in params]).cpu().numpy().tobytes(),)
        return key

    def refresh_engine(self):
        """Drop the packed weights on the device; the next generate call re-reads the parameters."""
        eng = self.__dict__.pop("_engine", None)
        self.__dict__.pop("_engine_ver", None)
        if eng is not None:
            eng.close()

    def _engine_for(self, device):
        ver = self._weights_version()
        eng = self.__dict__.get("_engine")
        if eng is None or self.__dict__.get("_engine_ver") != ver or eng.device != torch.device(device):
            if eng is not None:
                eng.close()
            eng = Engine(self._kind, self._cfg_kwargs(), self.state_dict(), device)
            self.__dict__["_engine"] = eng
            self.__dict__["_engine_ver"] = ver
        return eng

    def engine(self):
        """The live C-ABI handle wrapper (built on first use after ``.cuda()`` / weight changes)."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("%s: the generation path runs on a B200 only; call model.cuda() first "
                               "(no CPU fallback exists)" % type(self).__name__)
        return self._engine_for(dev)

    def _forward_engine(self, aux, audio, do, clip=False):
        """Shared part of the two ``forward`` methods: conditioning + swn_forward on the model's device."""
        if do and self.do_prob > 0:
            raise NotImplementedError("dropout (do=True) is a training-only feature; this forward is inference only")
        eng = self.engine()
        aux = torch.as_tensor(aux).to(eng.device, torch.float32).contiguous()
        if aux.dim() != 3 or aux.shape[1] != self.n_aux:
            raise ValueError("aux must be [B, n_aux, T]; got %s" % (tuple(aux.shape),))
        eng.condition(aux)
        return eng.forward(audio, clip=clip)

    # ------------------------------------------------------------------ the generate call
    def _generate(self, audio, aux, n_samples_list, intervals, noise, seed, forced, return_heads, batch_tile,
                  kernel_sel=0):
        eng = self.engine()
        dev = eng.device
        n_samples_list = [int(n) for n in n_samples_list]
        max_samples = max(n_samples_list)
        aux = torch.as_tensor(aux).to(dev, torch.float32).contiguous()
        if aux.dim() != 3 or aux.shape[1] != self.n_aux or aux.shape[0] != len(n_samples_list):
            raise ValueError("aux must be [B, n_aux, T] with B == len(n_samples_list); got %s" % (tuple(aux.shape),))
        seed_audio = None
        if audio is not None:
            audio = torch.as_tensor(audio)
            default = 0 if self._kind == "cswnv" else self.n_quantize // 2
            if bool((audio != default).any()):
                seed_audio = audio.reshape(aux.shape[0], -1)
        if self._kind == "dswnv" and forced is not None:
            f = torch.as_tensor(forced)
            if f.numel() and (int(f.min()) < 0 or int(f.max()) >= self.n_quantize):
                # the reference's OneHot scatter raises on such an index (dswnv.py:92)
                raise ValueError("forced class indices must lie in [0, %d)" % self.n_quantize)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())      # follows torch.manual_seed like the reference
        start = time.time()
        eng.condition(aux)
        out, heads = eng.generate(max_samples, seed=seed, noise=noise, forced=forced, seed_audio=seed_audio,
                                  return_heads=return_heads, batch_tile=batch_tile, kernel_sel=kernel_sel,
                                  n_samples_list=n_samples_list)
        samples = out.cpu().numpy()                                  # the reference's single D2H (cswnv:426)
        dt = time.time() - start
        n_gen = samples.shape[1]
        logging.info("average throughput / sample = %.9f sec (%ld samples * %ld) [%.3f kHz/s]" % (
            dt / max(n_gen * len(n_samples_list), 1), n_gen, len(n_samples_list),
            n_gen * len(n_samples_list) / (1000 * max(dt, 1e-12))))
        samples_list = [samples[i, :n] for i, n in enumerate(n_samples_list)]
        if return_heads:
            return samples_list, heads.cpu().numpy()
        return samples_list
