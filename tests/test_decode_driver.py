"""Decode driver.  Pinned by tests/golden_decode/*.npz: the reference's own decode scripts run
end to end (main(): model.conf + checkpoint load, n_gpus split, decode_generator, batch_fast_generate, clip /
decode_mu_law, sf.write) by tests/golden/make_decode_golden.py.  CPU: batching / padding / seeds / file formats;
GPU: ``decode.main`` on the same files and noise writes the same PCM."""
import argparse
import os
import wave

import numpy as np
import pytest
import torch

from _util import ROOT, load_golden
import shallow_wavenet_b200 as S
from shallow_wavenet_b200 import decode as D, sharding
from oracle import decode_oracle as DO


def _feats(n, n_aux, seed):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, 9, n)
    lens[n // 2] = lens[0]                       # ties in the length sort
    return {"/data/utt%02d.h5" % i: rng.standard_normal((int(t), n_aux)) for i, t in enumerate(lens)}


@pytest.mark.parametrize("kind,n,bs,seg", [("cswnv", 7, 3, 5), ("cswnv", 1, 7, 1), ("dswnv", 10, 4, 1), ("cswnv", 9, 9, 2), ("dswnv", 5, 1, 1)])
def test_decode_generator_matches_reference_batching(kind, n, bs, seg):
    feats = _feats(n, 6, n)
    flist = list(feats)
    got = list(D.decode_generator(flist, batch_size=bs, seg=seg, upsampling_factor=10, kind=kind,
                                  reader=lambda f, p: feats[f]))
    ref = list(DO.decode_generator(feats, flist, bs, seg, 10, kind))
    assert len(got) == len(ref)
    for (ids, (bx, bh, ns)), (rids, rx, rh, rns) in zip(got, ref):
        assert ids == rids and ns == rns
        assert bx.shape == tuple(rx.shape) and np.array_equal(bx.numpy(), rx)
        assert bx.dtype == (torch.float32 if kind == "cswnv" else torch.int64)
        assert bh.dtype == torch.float32 and np.array_equal(bh.numpy(), rh)
    assert sorted(i for ids, _ in got for i in ids) == sorted(os.path.basename(f)[:-3] for f in flist)


def test_mu_law_matches_reference_vectors():
    z = np.load(os.path.join(ROOT, "tests", "golden_aux", "mulaw.npz"))
    assert np.array_equal(S.encode_mu_law(z["x"], 256), z["idx256"])
    assert np.array_equal(S.encode_mu_law(z["x"], 1024), z["idx1024"])
    assert np.array_equal(S.decode_mu_law(z["y"], 256), z["wav256"])
    assert np.array_equal(S.decode_mu_law(np.arange(1024), 1024), z["wav1024"])
    assert int(S.encode_mu_law(np.zeros(1), 256)[0]) == 128     # the decode scripts' seed (decode_dswnv_softmax.py:94-99)


def test_pcm16_wav_files(tmp_path):
    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 0.99999, 1e-5, -3.0517578125e-05])
    assert D.pcm16(x).tolist() == [0, 16384, -16384, 32767, -32767, 32767, 0, -1]      # lrint(x * 0x7FFF), half to even
    p = str(tmp_path / "a.wav")
    D.write_wav(p, x, 22050)
    with wave.open(p, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 22050, len(x))
        assert np.array_equal(np.frombuffer(w.readframes(len(x)), np.int16), D.pcm16(x))



DECODE_GOLDEN = os.path.join(ROOT, "tests", "golden_decode")


def _load_decode_golden(name):
    z = np.load(os.path.join(DECODE_GOLDEN, name + ".npz"))
    conf = {k[5:]: z[k].item() for k in z.files if k.startswith("conf.")}
    return dict(kind=str(z["kind"]), fs=int(z["fs"]), batch_size=int(z["batch_size"]), n_gpus=int(z["n_gpus"]), conf=conf,
                order=[str(x) for x in z["order"]], batches=[str(b).split("|") for b in z["batches"]],
                sd={k[3:]: z[k] for k in z.files if k.startswith("sd.")},
                feats={k[5:]: z[k] for k in z.files if k.startswith("feat.")},
                noise={k[6:]: z[k] for k in z.files if k.startswith("noise.")},
                wav={k[4:]: z[k] for k in z.files if k.startswith("wav.")})


def _write_reference_format_files(g, d, with_h5_names=False):
    """model.conf (pickled Namespace, train_cswnv...py:293), checkpoint-7.pkl (save_checkpoint layout, :171-176), one
    feature file per utterance, and the noise archive for decode.main's test hook."""
    torch.save(argparse.Namespace(**g["conf"]), os.path.join(d, "model.conf"))
    torch.save({"model": {k: torch.from_numpy(v) for k, v in g["sd"].items()}, "optimizer": {"state": {}, "param_groups": []},
                "numpy_random_state": np.random.get_state(), "torch_random_state": torch.get_rng_state(), "iterations": 7},
               os.path.join(d, "checkpoint-7.pkl"))
    fdir = os.path.join(d, "feats")
    os.makedirs(fdir)
    for fid, a in g["feats"].items():
        np.save(os.path.join(fdir, fid + ".npy"), a)
    np.savez(os.path.join(d, "noise.npz"), **{"noise." + k: v for k, v in g["noise"].items()})
    return fdir


@pytest.mark.parametrize("name", ["decode_cswnv", "decode_dswnv"])
def test_batching_matches_the_reference_scripts_run(name):
    """Lists per GPU, batch composition and order, seeds, padding: as recorded from the reference's main()."""
    g = _load_decode_golden(name)
    files = sorted("/data/%s.h5" % f for f in g["feats"])                          # sorted(find_files(...)) (:191)
    lists = sharding.split_for_gpus(files, g["n_gpus"])                            # :200-201
    got = []
    U = g["conf"]["upsampling_factor"]
    for fl in lists:
        for ids, (bx, bh, ns) in D.decode_generator(fl, batch_size=g["batch_size"], seg=g["conf"].get("seg", 1),
                                                     upsampling_factor=U, kind=g["kind"],
                                                     reader=lambda f, p: g["feats"][os.path.basename(f)[:-3]]):
            got.append(ids)
            assert ns == [g["feats"][i].shape[0] * U for i in ids]
            assert bh.shape == (len(ids), g["conf"]["n_aux"], max(g["feats"][i].shape[0] for i in ids))
            for r, i in enumerate(ids):
                t = g["feats"][i].shape[0]
                assert np.array_equal(bh[r, :, :t].numpy(), g["feats"][i].T) and float(bh[r, :, t:].abs().sum()) == 0.0
            if g["kind"] == "dswnv":
                assert bx.dtype == torch.int64 and bx.tolist() == [[128]] * len(ids)
            else:
                assert bx.dtype == torch.float32 and bx.shape == (len(ids), g["conf"]["seg"]) and float(bx.abs().sum()) == 0.0
    assert got == g["batches"]
    assert [i for b in got for i in b] == g["order"]                                # files are written in this order
    feats_by_file = {f: g["feats"][os.path.basename(f)[:-3]] for f in files}
    rest = [ids for fl in lists for ids, _, _, _ in DO.decode_generator(feats_by_file, fl, g["batch_size"],
                                                                        g["conf"].get("seg", 1), U, g["kind"])]
    assert rest == g["batches"]                                                     # the oracle restatement agrees too


@pytest.mark.parametrize("name", ["decode_cswnv", "decode_dswnv"])
def test_reference_format_config_and_checkpoint_load(name, tmp_path):
    g = _load_decode_golden(name)
    fdir = _write_reference_format_files(g, str(tmp_path))
    config = D.load_config(str(tmp_path / "model.conf"))
    assert isinstance(config, argparse.Namespace) and config.string_path == "/feat_org_lf0"
    model = D.build_model(config)
    assert model._kind == g["kind"]
    model.load_state_dict(D.load_checkpoint_state(str(tmp_path / "checkpoint-7.pkl")), strict=True)
    assert D.find_feats(fdir) == sorted(os.path.join(fdir, f + ".npy") for f in g["feats"])
    lst = tmp_path / "feats.scp"
    lst.write_text("\n".join(D.find_feats(fdir)) + "\n")
    assert D.find_feats(str(lst)) == D.find_feats(fdir)
    # the outputs the reference wrote are what its own clip / decode_mu_law produce
    for fid, w in g["wav"].items():
        assert w.shape == ((g["feats"][fid].shape[0] * g["conf"]["upsampling_factor"]),)
        assert np.abs(w).max() <= 1.0 + (0.03 if g["kind"] == "dswnv" else 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["decode_cswnv", "decode_dswnv"])
def test_decode_main_writes_what_the_reference_script_wrote(name, tmp_path):
    """decode.main on reference-format files (model.conf, checkpoint-7.pkl), the reference's n_gpus / batch_size and the
    noise its sampler consumed: the PCM16 files equal the reference's sf.write input quantised the libsndfile way
    (CSWNV: free-running fp32 feedback, <= 5e-5 -> at most 2 LSB; DSWNV: bit-exact indices -> identical PCM)."""
    g = _load_decode_golden(name)
    fdir = _write_reference_format_files(g, str(tmp_path))
    out = str(tmp_path / "out")
    rc = D.main(["--feats", fdir, "--checkpoint", str(tmp_path / "checkpoint-7.pkl"), "--config", str(tmp_path / "model.conf"),
                 "--outdir", out, "--fs", str(g["fs"]), "--batch_size", str(g["batch_size"]), "--n_gpus", str(g["n_gpus"]),
                 "--verbose", "0"], noise_npz=str(tmp_path / "noise.npz"))
    assert rc == 0
    assert sorted(os.listdir(out)) == sorted([f + ".wav" for f in g["wav"]] + ["decode.log"])
    for fid, ref in g["wav"].items():
        with wave.open(os.path.join(out, fid + ".wav"), "rb") as w:
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, g["fs"], len(ref))
            pcm = np.frombuffer(w.readframes(len(ref)), np.int16).astype(np.int32)
        want = D.pcm16(ref).astype(np.int32)
        if g["kind"] == "dswnv":
            assert np.array_equal(pcm, want), fid
        else:
            assert int(np.abs(pcm - want).max()) <= 2, (fid, int(np.abs(pcm - want).max()))
