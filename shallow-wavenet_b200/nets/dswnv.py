# -*- coding: utf-8 -*-
"""B200-native drop-in for the generation path of the reference's discrete shallow WaveNet
vocoder (256-way mu-law softmax output).

Boundary mirrored: ``DSWNV.__init__`` (reference src/nets/dswnv.py:191-248), its
``state_dict`` layout, ``DSWNV.batch_fast_generate`` (:296-399) as called by
src/bin/decode_dswnv_softmax.py:216-255, and the mu-law helpers the decode script imports
(:19-47).
"""
import numpy as np

from ._base import _SWNBase, _ParamConv, initialize  # noqa: F401


def encode_mu_law(x, mu=256):
    """Mu-law companding + quantisation to [0, mu-1] (reference dswnv.py:19-31)."""
    m = mu - 1
    fx = np.sign(x) * np.log(1 + m * np.abs(x)) / np.log(1 + m)
    return np.floor((fx + 1) / 2 * m + 0.5).astype(np.int64)


def decode_mu_law(y, mu=256):
    """Inverse of ``encode_mu_law`` with the reference's half-step offset (dswnv.py:34-47)."""
    m = mu - 1
    fx = (y - 0.5) / m * 2 - 1
    return np.sign(fx) / m * ((1 + m) ** np.abs(fx) - 1)


class DSWNV(_SWNBase):
    _kind = "dswnv"

    def __init__(self, n_quantize=256, n_aux=54, hid_chn=192, skip_chn=256, aux_kernel_size=3,
                 aux_dilation_size=2, dilation_depth=3, dilation_repeat=3, kernel_size=6,
                 upsampling_factor=110, audio_in_flag=False, wav_conv_flag=False, do_prob=0):
        super(DSWNV, self).__init__()
        self.n_aux = n_aux
        self.n_quantize = n_quantize
        self.upsampling_factor = upsampling_factor
        self.in_audio_dim = n_quantize
        self.n_hidch = hid_chn
        self.n_skipch = skip_chn
        self.kernel_size = kernel_size
        self.dilation_depth = dilation_depth
        self.dilation_repeat = dilation_repeat
        self.aux_kernel_size = aux_kernel_size
        self.aux_dilation_size = aux_dilation_size
        self.do_prob = do_prob
        self.audio_in_flag = audio_in_flag
        self.wav_conv_flag = wav_conv_flag

        self._build_aux()
        self.in_aux_dim = n_aux * self.conv_aux.rec_field
        self.in_tot_dim = self.in_aux_dim + (self.in_audio_dim if audio_in_flag else 0)
        self._build_audio_in(in_audio_ch=n_quantize)
        self._build_cells(self.in_tot_dim)
        self.out_1 = _ParamConv((n_quantize, skip_chn, 1))
        self.out_2 = _ParamConv((n_quantize, n_quantize, 1))

    def _cfg_kwargs(self):
        return dict(n_quantize=self.n_quantize, n_aux=self.n_aux, hid_chn=self.n_hidch, skip_chn=self.n_skipch,
                    aux_kernel_size=self.aux_kernel_size, aux_dilation_size=self.aux_dilation_size,
                    dilation_depth=self.dilation_depth, dilation_repeat=self.dilation_repeat,
                    kernel_size=self.kernel_size, upsampling_factor=self.upsampling_factor, seg=1, lpc=0,
                    aux_conv2d_flag=0, wav_conv_flag=self.wav_conv_flag, audio_in_flag=self.audio_in_flag)

    def forward(self, audio, aux, do=False, last=False):
        """``DSWNV.forward`` of the reference (dswnv.py:250-276), inference only: ``audio`` is the one-hot float tensor
        [B, n_quantize, T_a] (or int64 class indices [B, T_a]), ``aux`` [B, n_aux, T_f] with T_a = T_f * U - 1; returns
        the logits [B, T_a, n_quantize] as a CUDA tensor.  Computed time-parallel by tcgen05 GEMMs (swn_forward)."""
        return self._forward_engine(aux, audio, do)["out"]

    def batch_fast_generate(self, audio, aux, n_samples_list, intervals=4410, *, noise=None, seed=None,
                            forced=None, return_heads=False, batch_tile=0, kernel_sel=0):
        """Same call as the reference (dswnv.py:296): ``audio`` [B, 1] int64 seed (mu-law index),
        ``aux`` [B, n_aux, T], ``n_samples_list`` lengths; returns a list of int64 numpy arrays of
        mu-law indices.

        Keyword-only extras: ``noise`` [B, N, 256] Exp(1) variates (the ones torch.multinomial
        would draw), ``seed`` Philox key, ``forced`` [B, N] indices fed back instead of the sampled
        ones, ``return_heads`` to also get the logits per step, ``batch_tile``, ``kernel_sel`` (0 auto,
        1 layer-group kernel, 2 cluster kernel).
        """
        return self._generate(audio, aux, n_samples_list, intervals, noise, seed, forced, return_heads, batch_tile, kernel_sel)
