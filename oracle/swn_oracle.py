# -*- coding: utf-8 -*-
"""CPU oracle for the shallow-wavenet autoregressive generation path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``shallow-wavenet_b200/`` may import
this file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker / the reported CPU baseline -- never as the thing shipped.

Parity status: PINNED.  The reference publishes no golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
modules themselves (``/root/reference/src/nets/{cswnv_shift1,dswnv}.py``
imported read-only in the build container by ``tests/golden/make_golden.py``);
the resulting fixtures live in ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks every function below against them.

It is a restatement (torch-CPU tensor ops + numpy, fp32) of what the reference
computes, written against a plain ``dict`` of weights named like the
reference's ``state_dict``:

* ``cond_features``        <- cswnv_shift1.py:297-311 / dswnv.py:302-309
* ``teacher_forced``       <- the full-sequence evaluation of the layers
                              (cswnv_shift1.py:314-334,367 / dswnv.py:312-331,361)
                              at the time indices the generation loop visits
* ``generate``             <- the hot loop cswnv_shift1.py:348-402 /
                              dswnv.py:338-374 (fixed-size history windows
                              instead of the reference's growing ``torch.cat``)
* ``forward_full``         <- the training-time forward, cswnv_shift1.py:191-267 / dswnv.py:250-276
* ``laplace_sample`` / ``softmax_sample`` <- cswnv_shift1.py:368-391 / dswnv.py:363-369
* ``encode_mu_law`` / ``decode_mu_law``    <- dswnv.py:19-47
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------
@dataclass
class SWNConfig:
    """Constructor arguments of the two vocoders (cswnv_shift1.py:131-133,
    dswnv.py:191-193)."""
    kind: str = "cswnv"            # "cswnv" (Laplace/LPC) | "dswnv" (mu-law softmax)
    n_quantize: int = 256
    n_aux: int = 54
    hid_chn: int = 192
    skip_chn: int = 256
    aux_kernel_size: int = 3
    aux_dilation_size: int = 2
    dilation_depth: int = 3
    dilation_repeat: int = 2
    kernel_size: int = 7
    upsampling_factor: int = 110
    seg: int = 1
    lpc: int = 0
    aux_conv2d_flag: bool = False
    wav_conv_flag: bool = False
    audio_in_flag: bool = False

    # derived quantities (cswnv_shift1.py:158-182, dswnv.py:214-241) ----------
    @property
    def dil_facts(self) -> List[int]:
        return list(range(self.dilation_depth)) * self.dilation_repeat

    @property
    def dilations(self) -> List[int]:
        return [self.kernel_size ** d for d in self.dil_facts]

    @property
    def padding(self) -> List[int]:
        k = self.kernel_size
        return [k ** (d + 1) - k ** d for d in self.dil_facts]

    @property
    def receptive_field(self) -> int:
        return sum(self.padding) + self.kernel_size - 1

    @property
    def aux_chn(self) -> int:
        return self.n_aux * self.aux_kernel_size ** self.aux_dilation_size

    @property
    def eff_seg(self) -> int:
        return self.seg if self.kind == "cswnv" else 1

    @property
    def head_dim(self) -> int:
        return 2 * self.seg + self.lpc if self.kind == "cswnv" else self.n_quantize

    def to_dict(self):
        return asdict(self)


def _t(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.detach().to(torch.float32).cpu()
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.float32)


def _weights(sd: Dict) -> Dict[str, torch.Tensor]:
    return {k: _t(v) for k, v in sd.items()}


# --------------------------------------------------------------------------
# mu-law (dswnv.py:19-47)
# --------------------------------------------------------------------------
def encode_mu_law(x, mu=256):
    mu = mu - 1
    fx = np.sign(x) * np.log(1 + mu * np.abs(x)) / np.log(1 + mu)
    return np.floor((fx + 1) / 2 * mu + 0.5).astype(np.int64)


def decode_mu_law(y, mu=256):
    mu = mu - 1
    fx = (y - 0.5) / mu * 2 - 1
    return np.sign(fx) / mu * ((1 + mu) ** np.abs(fx) - 1)


# --------------------------------------------------------------------------
# conditioning (cswnv_shift1.py:297-311, dswnv.py:302-309)
# --------------------------------------------------------------------------
def aux_frames(cfg: SWNConfig, W, aux: torch.Tensor) -> torch.Tensor:
    """scale_in -> conv_aux at frame rate: [B, n_aux, T_f] -> [B, C_a, T_f]."""
    x = F.conv1d(aux, W["scale_in.weight"], W["scale_in.bias"])
    ks = cfg.aux_kernel_size
    for i in range(cfg.aux_dilation_size):       # cswnv_shift1.py:105-113,124-127
        x = F.conv1d(x, W["conv_aux.conv.%d.weight" % i], W["conv_aux.conv.%d.bias" % i],
                     dilation=ks ** i, padding=(ks ** (i + 1) - ks ** i) // 2)
    return x


def upsample(cfg: SWNConfig, W, y: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d(1,1,(1,U),stride (1,U)) (cswnv_shift1.py:48-65):
    x[c, f*U + j] = y[c, f] * w[j] + b."""
    U = cfg.upsampling_factor
    w = W["upsampling.conv.weight"].reshape(U)
    b = W["upsampling.conv.bias"].reshape(())
    B, C, T = y.shape
    return (y.unsqueeze(-1) * w.view(1, 1, 1, U) + b).reshape(B, C, T * U)


def cond_features(cfg: SWNConfig, W, aux: torch.Tensor) -> torch.Tensor:
    """The per-time-step conditioning x' the layers consume, left padded by the
    receptive field with 'replicate' and (seg>1) seg-window stacked:
    -> [B, in_x in-channels, R + T_f*U - seg + 1]."""
    x = upsample(cfg, W, aux_frames(cfg, W, aux))
    R = cfg.receptive_field
    if R > 0:
        x = F.pad(x, (R, 0), "replicate")
    seg = cfg.eff_seg
    if cfg.kind == "cswnv" and seg > 1:
        if cfg.aux_conv2d_flag:                   # cswnv_shift1.py:306-308
            x = F.conv2d(x.unfold(2, seg, 1).permute(0, 1, 3, 2),
                         W["aux_conv2d.weight"], W["aux_conv2d.bias"]).squeeze(2)
        else:                                     # cswnv_shift1.py:309-311
            x = x.unfold(2, seg, 1).permute(0, 2, 1, 3)
            x = x.reshape(x.size(0), x.size(1), -1).permute(0, 2, 1)
    return x.contiguous()


# --------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------
def _causal_conv(x, w, b, dilation):
    """CausalConv1d.forward (cswnv_shift1.py:68-92): left zero pad, crop."""
    k = w.shape[2]
    pad = (k - 1) * dilation
    return F.conv1d(x, w, b, padding=pad, dilation=dilation)[:, :, : x.shape[2]]


def _cell(cfg, W, l, x, h):
    """_dcrnn_forward (cswnv_shift1.py:275-279 / dswnv.py:284-288)."""
    H = cfg.hid_chn
    g = F.conv1d(x, W["in_x.%d.weight" % l], W["in_x.%d.bias" % l]) * \
        _causal_conv(h, W["dil_h.%d.conv.weight" % l], W["dil_h.%d.conv.bias" % l], cfg.dilations[l])
    z = torch.sigmoid(g[:, :H])
    h = (1 - z) * torch.tanh(g[:, H:]) + z * h
    skip = F.conv1d(h, W["out_skip.%d.weight" % l], W["out_skip.%d.bias" % l])
    return skip, h


def _head(W, s):
    return F.conv1d(F.relu(F.conv1d(F.relu(s), W["out_1.weight"], W["out_1.bias"])),
                    W["out_2.weight"], W["out_2.bias"])


def _one_hot(idx: torch.Tensor, depth: int) -> torch.Tensor:
    """OneHot (dswnv.py:68-93) followed by the transpose at dswnv.py:312: [B,T] -> [B,depth,T]."""
    return F.one_hot(idx % depth, depth).to(torch.float32).transpose(1, 2)


def _audio_full(cfg: SWNConfig, gen, B: int, n: int):
    """The padded audio the layers see when ``gen`` [B, n] has been generated:
    CSWNV zeros(R + seg) ++ gen (float); DSWNV [Q//2]*(R+1) ++ gen (long)."""
    R = cfg.receptive_field
    if cfg.kind == "cswnv":
        a = torch.zeros(B, R + cfg.seg + n, dtype=torch.float32)
        if n:
            a[:, R + cfg.seg:] = _t(gen)[:, :n]
    else:
        a = torch.full((B, R + 1 + n), cfg.n_quantize // 2, dtype=torch.int64)
        if n:
            a[:, R + 1:] = torch.as_tensor(np.asarray(gen)).to(torch.int64)[:, :n]
    return a


def _input_hidden(cfg, W, a_full):
    """softsign(causal(...)) over the whole padded audio (cswnv_shift1.py:314-323,
    dswnv.py:312-320).  Returns hidden [B, H, T] (T = len - (seg-1)) and, for
    audio_in_flag, the one-hot audio to be concatenated to the conditioning."""
    extra = None
    if cfg.kind == "cswnv":
        u = a_full.unsqueeze(1)
        if cfg.wav_conv_flag:
            u = F.conv1d(u, W["wav_conv.weight"], W["wav_conv.bias"])
        h = F.softsign(_causal_conv(u, W["causal.conv.weight"], W["causal.conv.bias"], 1)[:, :, cfg.seg - 1:])
    else:
        u = _one_hot(a_full, cfg.n_quantize)
        if cfg.audio_in_flag:
            extra = u
        if cfg.wav_conv_flag:
            u = F.conv1d(u, W["wav_conv.weight"], W["wav_conv.bias"])
        h = F.softsign(_causal_conv(u, W["causal.conv.weight"], W["causal.conv.bias"], 1))
    return h, extra


def teacher_forced(cfg: SWNConfig, sd: Dict, aux, gen) -> np.ndarray:
    """Head outputs the generation loop would see at every step if ``gen``
    ([B, N], float for CSWNV / int for DSWNV) were the samples generated so far.

    Evaluates the layers over the whole padded sequence (zero-padded causal
    convs, exactly as the warm-up at cswnv_shift1.py:323-334 / dswnv.py:320-331
    does for its R+1 columns) and reads the head at t = R + i*seg
    (SURVEY.md 3.1 "time-index contract").

    Returns [B, N//seg, head_dim]  (raw out_2 rows: CSWNV [mu(seg), pre-sigmoid
    scale(seg), lp coefs(lpc)], DSWNV logits).
    """
    W = _weights(sd)
    aux = _t(aux)
    B = aux.shape[0]
    seg = cfg.eff_seg
    gen_np = np.asarray(gen)
    n_iter = gen_np.shape[1] // seg
    R = cfg.receptive_field
    with torch.no_grad():
        x = cond_features(cfg, W, aux)
        a_full = _audio_full(cfg, gen_np, B, n_iter * seg)
        h, extra = _input_hidden(cfg, W, a_full)
        T = R + (n_iter - 1) * seg + 1                  # last column needed
        h = h[:, :, :T]
        x = x[:, :, :T]
        if extra is not None:
            x = torch.cat((x, extra[:, :, :T]), 1)
        s = None
        for l in range(len(cfg.dil_facts)):
            skip, h = _cell(cfg, W, l, x, h)
            s = skip if s is None else s + skip
        out = _head(W, s)[:, :, R::seg]                 # [B, head_dim, n_iter]
    return out.transpose(1, 2).contiguous().numpy()


def teacher_forced_tail(cfg: SWNConfig, sd: Dict, aux, gen, n_tail: int) -> np.ndarray:
    """Head outputs of the LAST ``n_tail`` loop steps only.  The receptive field is finite
    (R columns, cswnv_shift1.py:182), so evaluating the layers over the last R + n_tail*seg
    columns with zero history gives exactly the values of the full-length evaluation there;
    this is what lets the full-size (BASELINE.json) runs be checked in seconds.
    Returns [B, n_tail, head_dim]."""
    W = _weights(sd)
    aux = _t(aux)
    B = aux.shape[0]
    seg = cfg.eff_seg
    gen_np = np.asarray(gen)
    n_iter = gen_np.shape[1] // seg
    R = cfg.receptive_field
    n_tail = min(n_tail, n_iter)
    with torch.no_grad():
        x = cond_features(cfg, W, aux)
        a_full = _audio_full(cfg, gen_np, B, n_iter * seg)
        T = R + (n_iter - 1) * seg + 1
        # columns needed: [t_lo, T) with R extra columns of context for the layers and k-1 for the causal conv
        t_lo = max(0, T - ((n_tail - 1) * seg + 1) - R - cfg.kernel_size)
        h, extra = _input_hidden(cfg, W, a_full[:, t_lo: T + seg - 1] if cfg.kind == "cswnv" else a_full[:, t_lo:T])
        if t_lo > 0:
            # the first k-1 columns saw zero padding instead of real audio: drop them
            drop = cfg.kernel_size - 1
            h = h[:, :, drop:]
            t_lo += drop
            if extra is not None:
                extra = extra[:, :, drop:]
        xw = x[:, :, t_lo:T]
        if extra is not None:
            xw = torch.cat((xw, extra), 1)
        s = None
        for l in range(len(cfg.dil_facts)):
            skip, h = _cell(cfg, W, l, xw, h)
            s = skip if s is None else s + skip
        out = _head(W, s)                                # columns t_lo .. T-1
        first = R + (n_iter - n_tail) * seg - t_lo
        out = out[:, :, first::seg]
    return out.transpose(1, 2).contiguous().numpy()


def forward_full(cfg: SWNConfig, sd: Dict, aux, audio, clip: bool = False):
    """The training-time ``forward`` over whole utterances (cswnv_shift1.py:191-267, dswnv.py:250-276): zero conv
    history before the first sample (no warm-up), stride-1 outputs.

    CSWNV: ``audio`` float [B, T_a], T_a = T_f*U - seg -> dict(out [B,T,2seg+lpc], mu, b, log_b, a [, b_clip, log_b_clip]),
    T = T_a - seg + 1.  DSWNV: ``audio`` int [B, T_a] (class indices), T_a = T_f*U - 1 -> dict(out = logits [B,T_a,Q])."""
    W = _weights(sd)
    aux = _t(aux)
    seg = cfg.eff_seg
    with torch.no_grad():
        x = upsample(cfg, W, aux_frames(cfg, W, aux))[:, :, seg:]                  # cswnv:193 / dswnv:252
        if cfg.kind == "cswnv":
            if seg > 1:
                if cfg.aux_conv2d_flag:                                              # cswnv:196-198
                    x = F.conv2d(x.unfold(2, seg, 1).permute(0, 1, 3, 2), W["aux_conv2d.weight"], W["aux_conv2d.bias"]).squeeze(2)
                else:                                                                # cswnv:199-201
                    x = x.unfold(2, seg, 1).permute(0, 2, 1, 3)
                    x = x.reshape(x.size(0), x.size(1), -1).permute(0, 2, 1)
            u = _t(audio).unsqueeze(1)
            if cfg.wav_conv_flag:
                u = F.conv1d(u, W["wav_conv.weight"], W["wav_conv.bias"])
            h = F.softsign(_causal_conv(u, W["causal.conv.weight"], W["causal.conv.bias"], 1)[:, :, seg - 1:])   # cswnv:204-207
        else:
            u = _one_hot(torch.as_tensor(np.asarray(audio)).to(torch.int64), cfg.n_quantize)
            if cfg.audio_in_flag:
                x = torch.cat((x, u), 1)                                             # dswnv:254-255
            if cfg.wav_conv_flag:
                u = F.conv1d(u, W["wav_conv.weight"], W["wav_conv.bias"])
            h = F.softsign(_causal_conv(u, W["causal.conv.weight"], W["causal.conv.bias"], 1))            # dswnv:258-261
        s = None
        for l in range(len(cfg.dil_facts)):
            skip, h = _cell(cfg, W, l, x, h)
            s = skip if s is None else s + skip
        out = _head(W, s).transpose(1, 2).contiguous()
    res = {"out": out.numpy()}
    if cfg.kind == "cswnv":
        log_b = F.logsigmoid(out[:, :, seg:2 * seg])
        res.update(mu=out[:, :, :seg].numpy(), b=torch.exp(log_b).numpy(), log_b=log_b.numpy(), a=out[:, :, 2 * seg:].numpy())
        if clip:
            lc = torch.clamp(log_b, min=-14.162084148244246758816564788835)
            res.update(b_clip=torch.exp(lc).numpy(), log_b_clip=lc.numpy())
    return res


# --------------------------------------------------------------------------
# sampling
# --------------------------------------------------------------------------
def laplace_sample(cfg: SWNConfig, out: torch.Tensor, lp_buf: Optional[torch.Tensor],
                   eps: torch.Tensor):
    """cswnv_shift1.py:368-391.  out [B, 2seg+lpc], lp_buf [B, lpc] oldest->newest,
    eps [B, seg] in (-0.4999, 0.5).  Returns samples [B, seg] and the new lp_buf."""
    seg, lpc = cfg.seg, cfg.lpc
    mus = out[:, :seg]
    bs = torch.exp(F.logsigmoid(out[:, seg:2 * seg]))
    if lpc > 0:
        ass = out[:, 2 * seg:].flip(-1)
        res = []
        for j in range(seg):
            lp = torch.sum(ass * lp_buf, -1)
            e = eps[:, j]
            s = torch.clamp(lp + mus[:, j] - bs[:, j] * e.sign() * torch.log1p(-2 * e.abs()), min=-1, max=1)
            res.append(s)
            lp_buf = torch.cat((lp_buf[:, 1:], s[:, None]), 1)
        return torch.stack(res, 1), lp_buf
    s = torch.clamp(mus - bs * eps.sign() * torch.log1p(-2 * eps.abs()), min=-1, max=1)
    return s, lp_buf


def softmax_sample(logits: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """dswnv.py:363-369.  OneHotCategorical(softmax(logits)).sample() is
    torch.multinomial(p, 1, True) == argmax(p / q), q ~ Exp(1).
    logits, q: [B, Q] -> idx [B]."""
    p = F.softmax(logits, dim=-1)
    return torch.argmax(p / q, dim=-1)


def gumbel_argmax(logits: np.ndarray, q: np.ndarray) -> np.ndarray:
    """The form the CUDA head evaluates: argmax_k(logit_k - log q_k); equal to
    ``softmax_sample`` except on fp near-ties (first max wins in both)."""
    g = logits.astype(np.float32) - np.log(q.astype(np.float32))
    return np.argmax(g, axis=-1)


# --------------------------------------------------------------------------
# the hot loop (cswnv_shift1.py:337-402, dswnv.py:333-374)
# --------------------------------------------------------------------------
def generate(cfg: SWNConfig, sd: Dict, aux, n_samples: int, noise,
             forced=None, return_heads: bool = False, num_threads: Optional[int] = None):
    """Incremental generation exactly as the reference's loop: per step a causal
    conv over the last 2k+seg-2 inputs, six cells over [history ++ new] windows,
    head on the last column, sampling, history shift.

    noise: CSWNV eps [B, N] (one uniform in (-0.4999,0.5) per generated sample);
           DSWNV q [B, N, Q] (Exp(1) variates of the multinomial fast path).
    forced: optional [B, N] samples fed back instead of the sampled ones
            (teacher forcing); sampled values are still returned.
    Returns samples [B, N'] (N' = (N//seg)*seg; float32 / int64) and, if asked,
    head outputs [B, N'//seg, head_dim].
    """
    if num_threads is not None:
        torch.set_num_threads(num_threads)
    W = _weights(sd)
    aux = _t(aux)
    B = aux.shape[0]
    seg, k, R, H = cfg.eff_seg, cfg.kernel_size, cfg.receptive_field, cfg.hid_chn
    L = len(cfg.dil_facts)
    n_iter = n_samples // seg
    is_c = cfg.kind == "cswnv"
    with torch.no_grad():
        x = cond_features(cfg, W, aux)
        # ---- warm-up over the padded seed (cswnv_shift1.py:314-334 / dswnv.py:312-331)
        a0 = _audio_full(cfg, None, B, 0)
        h, extra = _input_hidden(cfg, W, a0)               # [B,H,R+1]
        x_ = x[:, :, : R + 1]
        if extra is not None:
            x_ = torch.cat((x_, extra), 1)
        buf_size = [cfg.padding[l + 1] if l < L - 1 else k - 1 for l in range(L)]
        bufs = []
        for l in range(L):
            _, h = _cell(cfg, W, l, x_, h)
            bufs.append(h[:, :, -buf_size[l] - seg: -seg].clone())
        # ---- input history window (last 2k+seg-2 columns of the conv input)
        win = 2 * k + seg - 2
        if is_c:
            u_hist = a0[:, seg - 1:].unsqueeze(1)          # [B,1,R+1]
            if cfg.wav_conv_flag:
                u_hist = F.conv1d(u_hist, W["wav_conv.weight"], W["wav_conv.bias"])
            lp_buf = torch.zeros(B, cfg.lpc) if cfg.lpc > 0 else None
            out_samples = torch.zeros(B, n_iter * seg)
            noise_t = _t(noise)
        else:
            u_hist = _one_hot(a0, cfg.n_quantize)
            last_onehot = u_hist[:, :, -1:]
            if cfg.wav_conv_flag:
                u_hist = F.conv1d(u_hist, W["wav_conv.weight"], W["wav_conv.bias"])
            out_samples = torch.zeros(B, n_iter, dtype=torch.int64)
            noise_t = _t(noise)
        u_hist = u_hist[:, :, -win:] if u_hist.shape[2] >= win else \
            F.pad(u_hist, (win - u_hist.shape[2], 0))      # zeros before t=0 = conv padding
        heads = torch.zeros(B, n_iter, cfg.head_dim) if return_heads else None
        forced_t = None if forced is None else torch.as_tensor(np.asarray(forced))
        wc, bc = W["causal.conv.weight"], W["causal.conv.bias"]

        for i in range(n_iter):
            n = R + 1 + i * seg                              # columns so far
            x_ = x[:, :, n - seg: n]
            if cfg.audio_in_flag and not is_c:
                x_ = torch.cat((x_, last_onehot), 1)
            # causal over the window, keep the last k+seg-1 columns (cswnv:352 / dswnv:345)
            h = F.softsign(F.conv1d(u_hist, wc, bc)[:, :, -(k + seg - 1):])
            s = None
            new_bufs = []
            for l in range(L):
                d = cfg.dilations[l]
                g = F.conv1d(x_, W["in_x.%d.weight" % l], W["in_x.%d.bias" % l]) * \
                    F.conv1d(h, W["dil_h.%d.conv.weight" % l], W["dil_h.%d.conv.bias" % l], dilation=d)[:, :, -seg:]
                z = torch.sigmoid(g[:, :H])
                hn = (1 - z) * torch.tanh(g[:, H:]) + z * h[:, :, -seg:]
                skip = F.conv1d(hn[:, :, -1:], W["out_skip.%d.weight" % l], W["out_skip.%d.bias" % l])
                s = skip if s is None else s + skip
                h = torch.cat((bufs[l], hn), 2)
                new_bufs.append(h[:, :, -buf_size[l]:])
            bufs = new_bufs
            out = _head(W, s)[:, :, -1]                      # [B, head_dim]
            if return_heads:
                heads[:, i] = out
            if is_c:
                eps = noise_t[:, i * seg:(i + 1) * seg]
                smp, new_lp = laplace_sample(cfg, out, lp_buf, eps)
                out_samples[:, i * seg:(i + 1) * seg] = smp
                fed = smp if forced_t is None else forced_t[:, i * seg:(i + 1) * seg].to(torch.float32)
                if cfg.lpc > 0:
                    if forced_t is None:
                        lp_buf = new_lp
                    else:
                        lp_buf = torch.cat((lp_buf, fed), 1)[:, -cfg.lpc:]
                u_new = fed.unsqueeze(1)
                if cfg.wav_conv_flag:
                    u_new = F.conv1d(u_new, W["wav_conv.weight"], W["wav_conv.bias"])
            else:
                idx = softmax_sample(out, noise_t[:, i])
                out_samples[:, i] = idx
                fed = idx if forced_t is None else forced_t[:, i].to(torch.int64)
                last_onehot = _one_hot(fed[:, None], cfg.n_quantize)
                u_new = last_onehot
                if cfg.wav_conv_flag:
                    u_new = F.conv1d(u_new, W["wav_conv.weight"], W["wav_conv.bias"])
            u_hist = torch.cat((u_hist, u_new), 2)[:, :, -win:]
    res = out_samples.numpy()
    if return_heads:
        return res, heads.numpy()
    return res


def split_outputs(samples: np.ndarray, n_samples_list: List[int]):
    """cswnv_shift1.py:426-428 / dswnv.py:395-397."""
    return [samples[i, :n] for i, n in enumerate(n_samples_list)]
