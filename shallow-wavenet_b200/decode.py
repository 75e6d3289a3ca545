# -*- coding: utf-8 -*-
"""Decode driver of the generation path: the part of
``src/bin/decode_cswnv_laplace-shift1.py`` / ``src/bin/decode_dswnv_softmax.py`` around the
``batch_fast_generate`` call - batching by length, zero padding of the aux features, the seed
waveform, per-utterance truncation, PCM16 wav files and the throughput log lines.

h5py / soundfile are optional (absent in this image): features can come from ``.npy`` files or an
in-memory mapping, wav files are written with the standard library when soundfile is missing."""
import logging
import os
import time
import wave

import numpy as np
import torch

from . import sharding
from .nets import encode_mu_law, decode_mu_law


def pad_list(batch_list, pad_value=0.0):
    """decode_cswnv_laplace-shift1.py:30-48: list of (T_i, C) -> (B, T_max, C), padded with ``pad_value``
    (the reference ignores the argument and pads with zeros; so does this for pad_value=0)."""
    maxlen = max(b.shape[0] for b in batch_list)
    out = np.full((len(batch_list), maxlen, batch_list[0].shape[-1]), pad_value, dtype=np.float64)
    for i, b in enumerate(batch_list):
        out[i, :b.shape[0]] = b
    return out


def default_reader(featfile, string_path):
    """[T, n_aux] aux features of one utterance: hdf5 dataset ``string_path`` (utils.read_hdf5) or a ``.npy`` file."""
    if featfile.endswith(".npy"):
        return np.load(featfile)
    import h5py                                   # optional dependency, as in the reference
    with h5py.File(featfile, "r") as f:
        return f[string_path][()]


def decode_generator(feat_list, batch_size=7, string_path="/feat_org_lf0", seg=1, upsampling_factor=0, spk_trg=None,
                     min_idx=None, kind="cswnv", n_quantize=256, reader=default_reader, device=None):
    """decode_generator of the reference scripts (decode_cswnv_laplace-shift1.py:51-113,
    decode_dswnv_softmax.py:54-117): sort the utterances by number of frames, cut them into
    ceil(n / batch_size) nearly equal batches, and per batch yield
    ``feat_ids, (batch_x, batch_h, n_samples_list)`` with the seed waveform ``batch_x`` (zeros [B, seg] float /
    mu-law code of 0.0 [B, 1] long), ``batch_h`` float [B, n_aux, T_max] zero padded, ``n_samples_list`` = T_i * U."""
    if spk_trg is not None:
        string_path = ("/feat_cvmcep_cycvae-%s-%s" % (min_idx, spk_trg)) if min_idx is not None else "/feat_cvmcep_cycvae-" + spk_trg
    logging.info(string_path)
    feats = {f: np.asarray(reader(f, string_path)) for f in feat_list}
    feat_list = list(feat_list)
    lengths = [feats[f].shape[0] for f in feat_list]
    for idx in sharding.decode_batches(lengths, batch_size):
        batch = [feat_list[i] for i in idx]
        batch_h = pad_list([feats[f] for f in batch])
        n_samples_list = [feats[f].shape[0] * upsampling_factor for f in batch]
        feat_ids = [os.path.basename(f).replace(".h5", "").replace(".npy", "") for f in batch]
        if kind == "cswnv":
            batch_x = torch.zeros(len(batch), seg, dtype=torch.float32)
        else:
            batch_x = torch.full((len(batch), 1), int(encode_mu_law(np.zeros(1), n_quantize)[0]), dtype=torch.long)
        batch_h = torch.FloatTensor(batch_h).transpose(1, 2).contiguous()
        if device is not None:
            batch_x, batch_h = batch_x.to(device), batch_h.to(device)
        yield feat_ids, (batch_x, batch_h, n_samples_list)


def pcm16(wav):
    """float [-1, 1] -> int16 the way libsndfile writes PCM_16 from floating-point data with its defaults
    (normalisation on, clipping off: ``lrint(x * 0x7FFF)``, libsndfile pcm.c d2s_array) - what
    ``sf.write(path, wav, fs, "PCM_16")`` does in decode_cswnv_laplace-shift1.py:253.  The callers clip to [-1, 1]
    first, as the reference does (:252); the clip here only guards the int16 cast."""
    return np.clip(np.rint(np.asarray(wav, np.float64) * 32767.0), -32768, 32767).astype(np.int16)


def write_wav(path, wav, fs):
    """sf.write(path, wav, fs, "PCM_16") (decode_cswnv_laplace-shift1.py:253); standard library fallback."""
    try:
        import soundfile as sf
        sf.write(path, wav, fs, "PCM_16")
        return
    except ImportError:
        pass
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(fs))
        w.writeframes(pcm16(wav).tobytes())


def gpu_decode(model, feat_list, outdir, fs, batch_size=7, intervals=4410, string_path="/feat_org_lf0", spk_trg=None,
               min_idx=None, reader=default_reader, seed=None):
    """Body of ``gpu_decode`` (decode_cswnv_laplace-shift1.py:230-259, decode_dswnv_softmax.py:236-266) for a
    model that already holds its weights on the GPU.  Returns {feat_id: waveform}."""
    kind = model._kind
    os.makedirs(outdir, exist_ok=True)
    gen = decode_generator(feat_list, batch_size=batch_size, string_path=string_path, spk_trg=spk_trg, min_idx=min_idx,
                           seg=getattr(model, "seg", 1), upsampling_factor=model.upsampling_factor, kind=kind,
                           n_quantize=getattr(model, "n_quantize", 256), reader=reader)
    time_sample, n_samples, n_samples_t, out = [], [], [], {}
    for feat_ids, (batch_x, batch_h, n_samples_list) in gen:
        logging.info("decoding start")
        start = time.time()
        kw = {} if seed is None else {"seed": seed + len(out)}
        samples_list = model.batch_fast_generate(batch_x, batch_h, n_samples_list, intervals, **kw)
        time_sample.append(time.time() - start)
        n_samples.append(max(n_samples_list))
        n_samples_t.append(max(n_samples_list) * len(n_samples_list))
        for feat_id, samples in zip(feat_ids, samples_list):
            wav = np.clip(samples, -1, 1) if kind == "cswnv" else decode_mu_law(samples, model.n_quantize)
            write_wav(os.path.join(outdir, feat_id + ".wav"), wav, fs)
            logging.info("wrote %s.wav in %s." % (feat_id, outdir))
            out[feat_id] = wav
    if time_sample:
        logging.info("average time / sample = %.6f sec (%ld samples) [%.3f kHz/s]" % (
            sum(time_sample) / sum(n_samples), sum(n_samples), sum(n_samples) / (1000 * sum(time_sample))))
        logging.info("average throughput / sample = %.6f sec (%ld samples) [%.3f kHz/s]" % (
            sum(time_sample) / sum(n_samples_t), sum(n_samples_t), sum(n_samples_t) / (1000 * sum(time_sample))))
    return out
