# -*- coding: utf-8 -*-
"""B200-native drop-in for the generation path of the reference's continuous shallow
WaveNet vocoder (Laplacian output, multiple samples per step, data-driven linear prediction).

Boundary mirrored: ``CSWNV.__init__`` (reference src/nets/cswnv_shift1.py:131-189),
its ``state_dict`` layout, and ``CSWNV.batch_fast_generate`` (:287-430) as called by
src/bin/decode_cswnv_laplace-shift1.py:208-248.
"""
from torch import nn

import torch

from ._base import _SWNBase, _ParamConv, initialize  # noqa: F401


class CSWNV(_SWNBase):
    _kind = "cswnv"

    def __init__(self, n_aux=54, hid_chn=192, skip_chn=256, aux_kernel_size=3, aux_dilation_size=2,
                 dilation_depth=3, dilation_repeat=2, kernel_size=7, upsampling_factor=110, seg=5, lpc=4,
                 do_prob=0, aux_conv2d_flag=False, wav_conv_flag=False):
        super(CSWNV, self).__init__()
        self.n_aux = n_aux
        self.n_hidch = hid_chn
        self.n_skipch = skip_chn
        self.aux_kernel_size = aux_kernel_size
        self.aux_dilation_size = aux_dilation_size
        self.dilation_depth = dilation_depth
        self.dilation_repeat = dilation_repeat
        self.kernel_size = kernel_size
        self.upsampling_factor = upsampling_factor
        self.seg = seg
        self.lpc = lpc
        self.lpc_offset = seg - lpc
        self.do_prob = do_prob
        self.aux_conv2d_flag = aux_conv2d_flag
        self.wav_conv_flag = wav_conv_flag

        self._build_aux()
        self.in_aux_dim = n_aux * self.conv_aux.rec_field
        if aux_conv2d_flag and seg > 1:
            self.aux_conv2d = _ParamConv((self.in_aux_dim, self.in_aux_dim, seg, 1))
        elif seg > 1:
            self.in_aux_dim *= seg
        self._build_audio_in(in_audio_ch=1)
        self._build_cells(self.in_aux_dim)
        self.out_1 = _ParamConv((skip_chn, skip_chn, 1))
        self.out_2 = _ParamConv((2 * seg + lpc, skip_chn, 1))

    def _cfg_kwargs(self):
        return dict(n_quantize=0, n_aux=self.n_aux, hid_chn=self.n_hidch, skip_chn=self.n_skipch,
                    aux_kernel_size=self.aux_kernel_size, aux_dilation_size=self.aux_dilation_size,
                    dilation_depth=self.dilation_depth, dilation_repeat=self.dilation_repeat,
                    kernel_size=self.kernel_size, upsampling_factor=self.upsampling_factor, seg=self.seg,
                    lpc=self.lpc, aux_conv2d_flag=self.aux_conv2d_flag, wav_conv_flag=self.wav_conv_flag,
                    audio_in_flag=0)

    def forward(self, aux, audio, do=False, clip=False):
        """``CSWNV.forward`` of the reference (cswnv_shift1.py:191-267), inference only.  Argument order (aux, audio) as
        there: ``aux`` [B, n_aux, T_f], ``audio`` [B, 1, T_a] (or [B, T_a]) with T_a = T_f * U - seg.  Returns CUDA tensors in
        the reference's layouts: ``(mu, b, log_b, a)`` for lpc > 0 (mu / b / log_b [B, T, seg], a [B, T, lpc], T = T_a - seg + 1),
        ``(mu, b, log_b)`` without LP coefficients ([B, T] each when seg == 1); with ``clip=True`` the clipped scale pair
        is inserted after ``b`` as the reference does (:233-240).  Computed time-parallel by tcgen05 GEMMs (swn_forward)."""
        audio = torch.as_tensor(audio)
        if audio.dim() == 3:
            if audio.shape[1] != 1:
                raise ValueError("audio must be [B, 1, T]; got %s" % (tuple(audio.shape),))
            audio = audio[:, 0]
        r = self._forward_engine(aux, audio, do, clip)
        seg, lpc = self.seg, self.lpc
        out = r["out"]
        mu = out[:, :, :seg]
        scale = (r["b"], r["b_clip"], r["log_b_clip"]) if clip else (r["b"], r["log_b"])
        if lpc > 0:
            return (mu,) + scale + (out[:, :, 2 * seg:],)
        if seg > 1:
            return (mu,) + scale
        B = out.shape[0]
        return (mu.reshape(B, -1),) + tuple(x.reshape(B, -1) for x in scale)

    def batch_fast_generate(self, audio, aux, n_samples_list, intervals=4410, Laplace=True, *, noise=None,
                            seed=None, forced=None, return_heads=False, batch_tile=0, kernel_sel=0):
        """Same call as the reference (cswnv_shift1.py:287): ``audio`` [B, seg] seed waveform,
        ``aux`` [B, n_aux, T] zero-padded features, ``n_samples_list`` lengths; returns a list of
        float32 numpy arrays in [-1, 1].

        Keyword-only extras (not in the reference): ``noise`` [B, N] injected uniforms in
        (-0.4999, 0.5) replacing the Philox stream, ``seed`` the Philox key, ``forced`` [B, N]
        samples fed back instead of the sampled ones (teacher forcing), ``return_heads`` to also
        get the raw out_2 rows per step, ``batch_tile`` to pin the kernel's batch tile, ``kernel_sel`` (0 auto, 1 layer-group kernel,
        2 cluster kernel).
        """
        if not Laplace:
            # the reference returns an unbound local in this case (cswnv_shift1.py:343,423)
            raise ValueError("only the Laplacian output distribution exists on this path")
        return self._generate(audio, aux, n_samples_list, intervals, noise, seed, forced, return_heads, batch_tile, kernel_sel)
