# -*- coding: utf-8 -*-
"""Decode driver of the generation path: the part of
``src/bin/decode_cswnv_laplace-shift1.py`` / ``src/bin/decode_dswnv_softmax.py`` around the
``batch_fast_generate`` call - batching by length, zero padding of the aux features, the seed
waveform, per-utterance truncation, PCM16 wav files and the throughput log lines.

h5py / soundfile are optional (absent in this image): features can come from ``.npy`` files or an
in-memory mapping, wav files are written with the standard library when soundfile is missing."""
import argparse
import logging
import os
import sys
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
               min_idx=None, reader=default_reader, seed=None, noise_fn=None):
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
        if noise_fn is not None:                   # parity tests: the noise the reference's sampler consumed
            kw["noise"] = noise_fn(feat_ids, n_samples_list)
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


# ----------------------------------------------------------------------------- main() of the decode scripts
def load_config(path):
    """``config = torch.load(args.config)`` (decode_cswnv_laplace-shift1.py:187): ``model.conf`` is the pickled
    argparse.Namespace the training script saved (train_cswnv_laplace-stftcmplx_shift1.py:293); torch >= 2.6 only
    unpickles it with ``weights_only=False``."""
    return torch.load(path, map_location="cpu", weights_only=False)


def load_checkpoint_state(path):
    """``torch.load(args.checkpoint)["model"]`` (:224): ``checkpoint-<iter>.pkl`` holds
    {"model", "optimizer", "numpy_random_state", "torch_random_state", "iterations"} (train_...:171-176); the final
    checkpoint only {"model"} (:894-896)."""
    return torch.load(path, map_location="cpu", weights_only=False)["model"]


def build_model(config, kind=None):
    """The constructor calls of the two scripts (decode_cswnv_laplace-shift1.py:208-221, decode_dswnv_softmax.py:216-228);
    ``kind`` defaults to DSWNV when the config carries ``n_quantize``."""
    from .nets import CSWNV, DSWNV
    if kind is None:
        kind = "dswnv" if hasattr(config, "n_quantize") else "cswnv"
    if kind == "cswnv":
        return CSWNV(n_aux=config.n_aux, skip_chn=config.skip_chn, hid_chn=config.hid_chn,
                     dilation_depth=config.dilation_depth, dilation_repeat=config.dilation_repeat,
                     kernel_size=config.kernel_size, aux_kernel_size=config.aux_kernel_size,
                     aux_dilation_size=config.aux_dilation_size, seg=config.seg, lpc=config.lpc,
                     aux_conv2d_flag=config.aux_conv2d_flag, wav_conv_flag=config.wav_conv_flag,
                     upsampling_factor=config.upsampling_factor)
    return DSWNV(n_quantize=config.n_quantize, n_aux=config.n_aux, hid_chn=config.hid_chn, skip_chn=config.skip_chn,
                 dilation_depth=config.dilation_depth, dilation_repeat=config.dilation_repeat,
                 kernel_size=config.kernel_size, aux_kernel_size=config.aux_kernel_size,
                 aux_dilation_size=config.aux_dilation_size, audio_in_flag=config.audio_in,
                 wav_conv_flag=config.wav_conv_flag, upsampling_factor=config.upsampling_factor)


def find_feats(feats):
    """--feats: a directory searched recursively for *.h5 (sorted; *.npy accepted too) or a text list of files
    (decode_cswnv_laplace-shift1.py:190-197, utils.find_files / read_txt)."""
    if os.path.isdir(feats):
        found = []
        for root, _, files in os.walk(feats, followlinks=True):
            found += [os.path.join(root, f) for f in files if f.endswith((".h5", ".npy"))]
        return sorted(found)
    if os.path.isfile(feats):
        with open(feats, "r") as f:
            return [line.replace("\n", "") for line in f.readlines()]
    raise ValueError("--feats should be directory or list.")


class _NpzNoise(object):
    """Picklable noise source for the parity tests: an .npz with one ``noise.<feat_id>`` array per utterance."""

    def __init__(self, path):
        self.path = path

    def __call__(self, feat_ids, n_samples_list):
        z = np.load(self.path)
        n = max(n_samples_list)
        return np.stack([z["noise." + f][:n] for f in feat_ids])


def _decode_worker(gpu, feat_lists, opts):
    """One process per GPU (decode_cswnv_laplace-shift1.py:204-259); a top-level function so that it can be spawned
    (the reference forks a closure, which a CUDA-initialised parent cannot do safely)."""
    if opts.get("log_file"):
        _setup_logging(opts["verbose"], opts["log_file"])
    os.environ["PYTHONHASHSEED"] = str(opts["seed"])
    np.random.seed(opts["seed"])
    torch.manual_seed(opts["seed"])
    config = load_config(opts["config"])
    with torch.cuda.device(gpu), torch.no_grad():
        model = build_model(config, opts.get("kind"))
        logging.info(model)
        model.cuda()
        model.load_state_dict(load_checkpoint_state(opts["checkpoint"]))
        model.eval()
        for param in model.parameters():
            param.requires_grad = False
        noise_fn = _NpzNoise(opts["noise_npz"]) if opts.get("noise_npz") else None
        for feat_list in feat_lists:
            gpu_decode(model, feat_list, opts["outdir"], opts["fs"], batch_size=opts["batch_size"],
                       intervals=opts["intervals"], string_path=config.string_path, spk_trg=opts["spk_trg"],
                       min_idx=opts["min_idx"], noise_fn=noise_fn)


def _setup_logging(verbose, filename):
    level = logging.INFO if verbose > 0 else logging.WARN
    logging.basicConfig(level=level, format="%(asctime)s (%(module)s:%(lineno)d) %(levelname)s: %(message)s",
                        datefmt="%m/%d/%Y %I:%M:%S", filename=filename, force=True)
    logging.getLogger().addHandler(logging.StreamHandler())
    if verbose <= 0:
        logging.warning("logging is disabled.")


def main(argv=None, kind=None, noise_npz=None):
    """The command line of src/bin/decode_cswnv_laplace-shift1.py / decode_dswnv_softmax.py (same flags, :116-148):
    load ``model.conf`` and the checkpoint, split the list into ``n_gpus`` contiguous chunks (:200-201) and decode each
    chunk in its own process on its own GPU (:262-274).  Differences, none of them on the data path: the workers are
    spawned top-level functions instead of forked closures; with ``n_gpus == 1`` the single list is decoded in this
    process; when there are more lists than visible devices, list i goes to device ``i % device_count`` and the lists
    of one device are decoded one after the other by that device's process."""
    parser = argparse.ArgumentParser()
    parser.add_argument("--feats", required=True, type=str, help="list or directory of aux feat files")
    parser.add_argument("--checkpoint", required=True, type=str, help="model file")
    parser.add_argument("--config", required=True, type=str, help="configure file")
    parser.add_argument("--outdir", required=True, type=str, help="directory to save generated samples")
    parser.add_argument("--fs", default=22050, type=int, help="sampling rate")
    parser.add_argument("--batch_size", default=1, type=int, help="number of batch size in decoding")
    parser.add_argument("--n_gpus", default=1, type=int, help="number of gpus")
    parser.add_argument("--spk_trg", default=None, type=str)
    parser.add_argument("--min_idx", default=None, type=int)
    parser.add_argument("--intervals", default=4410, type=int, help="log interval")
    parser.add_argument("--seed", default=1, type=int, help="seed number")
    parser.add_argument("--GPU_device", default=None, type=int, help="selection of GPU device")
    parser.add_argument("--GPU_device_str", default=None, type=str, help="selection of GPU device")
    parser.add_argument("--verbose", default=1, type=int, help="log level")
    args = parser.parse_args(argv)

    # the reference always pins CUDA_VISIBLE_DEVICES (:150-154); here only when asked, so that a caller that has
    # already initialised CUDA (tests, bench) keeps its devices
    if args.GPU_device_str is not None or args.GPU_device is not None:
        os.environ["CUDA_DEVICE_ORDER"] = "PCI_BUS_ID"
        os.environ["CUDA_VISIBLE_DEVICES"] = args.GPU_device_str if args.GPU_device_str is not None else str(args.GPU_device)
    os.makedirs(args.outdir, exist_ok=True)
    log_file = os.path.join(args.outdir, "decode.log")
    _setup_logging(args.verbose, log_file)
    config = load_config(args.config)
    logging.info(config)
    feat_list = find_feats(args.feats)
    feat_lists = sharding.split_for_gpus(feat_list, args.n_gpus)                  # :200-201
    opts = dict(config=args.config, checkpoint=args.checkpoint, outdir=args.outdir, fs=args.fs, batch_size=args.batch_size,
                intervals=args.intervals, spk_trg=args.spk_trg, min_idx=args.min_idx, seed=args.seed, kind=kind,
                verbose=args.verbose, noise_npz=noise_npz, log_file=None)
    if args.n_gpus == 1:
        _decode_worker(torch.cuda.current_device() if torch.cuda.is_available() else 0, feat_lists, opts)
        return 0
    n_dev = max(1, torch.cuda.device_count())
    per_dev = {}
    for i, fl in enumerate(feat_lists):                                            # :262-271 (gpu = i mod n_gpus)
        per_dev.setdefault((i % args.n_gpus) % n_dev, []).append(fl)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    opts["log_file"] = log_file
    procs = [ctx.Process(target=_decode_worker, args=(gpu, lists, opts)) for gpu, lists in sorted(per_dev.items())]
    for p in procs:
        p.start()
    for p in procs:                                                                # :273-274
        p.join()
    bad = [p.exitcode for p in procs if p.exitcode != 0]
    if bad:
        raise RuntimeError("decode worker(s) exited with %s" % bad)
    return 0


if __name__ == "__main__":
    sys.exit(main())
