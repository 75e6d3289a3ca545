# -*- coding: utf-8 -*-
"""Noise-shaping restore filter (recipe stages 6 / 9) on the GPU: the host side of the reference's
``src/bin/noise_shaping.py`` - same function names, argument meaning and defaults - over the C ABI entry point
``swn_noise_shaping`` (hand-written CUDA, ``csrc/swn_post.cu``; float64 like the reference).

``synthesis_diff`` (``noise_shaping.py:52-87``) = pysptk ``mc2b`` + ``MLSADF(dim, alpha)`` + ``Synthesizer``;
``low_cut_filter`` (``feature_extract.py:57-76``) = ``lfilter(firwin(255, cutoff / (fs // 2), pass_zero=False), 1, x)``;
``noise_shaping`` = the per-file body of the script's worker (``:146-181``), batched: every waveform of the list is one
row of one launch.  There is no CPU path here: without the CUDA library and a B200 every call raises.
pysptk / pyworld / scipy are not needed: ``pyworld.harvest`` is only used by the script for its frame count
(``int(1000 * len / fs / frame_period) + 1``) and the FIR design is the closed form of scipy's ``firwin`` (Hamming window).
"""
import argparse
import ctypes
import os
import sys

import numpy as np

from . import _abi

FS = 24000
SHIFTMS = 5.0
FFTL = 1024
MCEP_DIM_START = 5
MCEP_ALPHA = 0.466
MAG = 0.5


def firwin_highpass(numtaps, cutoff_norm):
    """``scipy.signal.firwin(numtaps, cutoff_norm, pass_zero=False)`` (Hamming window, unit gain at Nyquist)."""
    if numtaps % 2 != 1:
        raise ValueError("a high-pass FIR needs an odd number of taps")
    m = np.arange(0, numtaps) - 0.5 * (numtaps - 1)
    h = (np.sinc(m) - cutoff_norm * np.sinc(cutoff_norm * m)) * np.hamming(numtaps)
    return h / np.sum(h * np.cos(np.pi * m))


def harvest_num_frames(n_samples, fs, frame_period=5.0):
    """Length of ``pyworld.harvest``'s f0 track - all the script uses of it (``noise_shaping.py:160-161``)."""
    return int(1000.0 * n_samples / fs / frame_period) + 1


def _run(xs, mcep, n_frames, alpha, hop, fir, pade=4, device=None):
    import torch
    if not torch.cuda.is_available():
        raise _abi.SwnError("the noise-shaping filter runs on a B200 only; there is no CPU fallback")
    lib = _abi.load()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    B = len(xs)
    lens = np.array([len(x) for x in xs], dtype=np.int64)
    stride = int(max(int(lens.max()), 1))
    host = np.zeros((B, stride), dtype=np.float64)
    for i, x in enumerate(xs):
        host[i, :len(x)] = np.asarray(x, dtype=np.float64)
    x_dev = torch.from_numpy(host).to(dev)
    y_dev = torch.empty_like(x_dev)
    tmp = torch.empty_like(x_dev) if fir is not None else None
    if mcep is None:                                           # the low-cut FIR alone
        utts, frames, dim = 1, 1, 1
    else:
        mcep = np.ascontiguousarray(mcep, dtype=np.float64)
        if mcep.ndim == 2:
            mcep = mcep[None]
        utts, frames, dim = mcep.shape
    nf = None if n_frames is None else np.ascontiguousarray(n_frames, dtype=np.int32)
    fir_a = None if fir is None else np.ascontiguousarray(fir, dtype=np.float64)
    a = _abi.SwnNsArgs(x_dev=x_dev.data_ptr(), y_dev=y_dev.data_ptr(), tmp_dev=tmp.data_ptr() if tmp is not None else None,
                       stride=stride, B=B, order=dim - 1, n_samples=lens.ctypes.data,
                       n_frames=nf.ctypes.data if nf is not None else None, mcep_host=mcep.ctypes.data if mcep is not None else None, mcep_utts=utts,
                       mcep_frames=frames, pade=pade, hop=int(hop), alpha=float(alpha),
                       fir_host=fir_a.ctypes.data if fir_a is not None else None, fir_taps=0 if fir_a is None else len(fir_a))
    _abi.check(lib.swn_noise_shaping(dev.index or 0, ctypes.byref(a), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    out = y_dev.cpu().numpy()
    return [out[i, :lens[i]].copy() for i in range(B)]


def synthesis_diff(x, diffmcep, rmcep=None, alpha=MCEP_ALPHA, fs=FS, shiftms=SHIFTMS):
    """filtering with a differential mel-cepstrum (``noise_shaping.py:52-87``): x (samples,), diffmcep (T, dim)."""
    if rmcep is not None:
        raise NotImplementedError("power modification (rmcep) is not on this path: noise_shaping.py's main() never passes it")
    shiftl = int(fs / 1000 * shiftms)
    return _run([x], np.asarray(diffmcep, dtype=np.float64), None, alpha, shiftl, None)[0]


def low_cut_filter(x, fs, cutoff=70):
    """``feature_extract.py:57-76`` on the GPU (the same entry point without a mel-cepstrum: the FIR alone)."""
    return _run([x], None, None, 0.0, 1, firwin_highpass(255, cutoff / (fs // 2)))[0]


def noise_shaping(wavs, mean_feat, fs=FS, shiftms=SHIFTMS, mcep_dim_start=MCEP_DIM_START, mcep_alpha=MCEP_ALPHA, mag=MAG,
                  inv=False):
    """The worker of ``noise_shaping.py:146-181`` for a list of waveforms (float arrays in [-1, 1]), one launch: mean
    mel-cepstrum * mag with c(0) = 0 (sign flipped for ``inv``), tiled over harvest's frame count, MLSA filter, 70 Hz low cut."""
    coef = np.array(mean_feat, dtype=np.float64)[mcep_dim_start:] * mag
    coef[0] = 0.0
    if inv:
        coef[1:] = -1.0 * coef[1:]
    frames = [harvest_num_frames(len(x), fs) for x in wavs]
    return _run(list(wavs), coef[None, :], frames, mcep_alpha, int(fs / 1000 * shiftms), firwin_highpass(255, 70 / (fs // 2)))


def _read_wav(path):
    try:
        import soundfile as sf
        return sf.read(path)
    except ImportError:
        import wave
        with wave.open(path, "rb") as w:
            if w.getsampwidth() != 2 or w.getnchannels() != 1:
                raise ValueError("%s: only mono PCM16 without soundfile" % path)
            x = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float64) / 32768.0   # libsndfile's scaling
            return x, w.getframerate()


def _read_mean(stats, key="/mean_org_lf0"):
    if stats.endswith(".npy"):
        return np.load(stats)
    import h5py
    with h5py.File(stats, "r") as f:
        return f[key][()]


def main(argv=None):
    """``noise_shaping.py:90-200`` with the same flags; the file list is filtered as ONE batch on the current GPU
    (``--n_jobs`` is accepted and ignored: the reference forks CPU workers)."""
    from .decode import write_wav
    ap = argparse.ArgumentParser(description="noise-shaping restore / apply filter on the GPU")
    ap.add_argument("--waveforms", required=True, help="directory or list file of wav files")
    ap.add_argument("--stats", required=True, help="hdf5 (needs h5py) or .npy with the mean feature vector")
    ap.add_argument("--writedir", required=True)
    ap.add_argument("--fs", default=FS, type=int)
    ap.add_argument("--shiftms", default=SHIFTMS, type=float)
    ap.add_argument("--fftl", default=FFTL, type=int)
    ap.add_argument("--mcep_dim_start", default=MCEP_DIM_START, type=int)
    ap.add_argument("--mcep_alpha", default=MCEP_ALPHA, type=float)
    ap.add_argument("--mag", default=MAG, type=float)
    ap.add_argument("--verbose", default=1, type=int)
    ap.add_argument("--n_jobs", default=1, type=int)
    ap.add_argument("--inv", default="false")
    args = ap.parse_args(argv)
    inv = str(args.inv).lower() in ("1", "true", "yes", "y", "t", "on")
    if os.path.isdir(args.waveforms):
        files = sorted(os.path.join(r, f) for r, _, fs_ in os.walk(args.waveforms) for f in fs_ if f.endswith(".wav"))
    else:
        files = [ln.strip() for ln in open(args.waveforms) if ln.strip()]
    os.makedirs(args.writedir, exist_ok=True)
    wavs = []
    for f in files:
        x, fs = _read_wav(f)
        if fs != args.fs:
            print("ERROR: sampling frequency is not matched.")
            sys.exit(1)
        wavs.append(x)
    outs = noise_shaping(wavs, _read_mean(args.stats), fs=args.fs, shiftms=args.shiftms, mcep_dim_start=args.mcep_dim_start,
                         mcep_alpha=args.mcep_alpha, mag=args.mag, inv=inv)
    for f, y in zip(files, outs):
        write_wav(os.path.join(args.writedir, os.path.basename(f)), y, args.fs)
    return 0


if __name__ == "__main__":
    sys.exit(main())
