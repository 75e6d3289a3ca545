"""Decode driver: batching / padding / seeds against the oracle restatement of the reference
scripts, mu-law maps against vectors generated from the reference module, PCM16 files, and (GPU) the whole
generate-and-write loop."""
import os
import wave

import numpy as np
import pytest
import torch

from _util import ROOT, load_golden
import shallow_wavenet_b200 as S
from shallow_wavenet_b200 import decode as D
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


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cswnv_seg5_lpc4_wav", "dswnv_base"])
def test_gpu_decode_writes_what_generate_returns(name, tmp_path):
    g = load_golden(name)
    kind, kw = g["kind"], g["kw"]
    m = (S.CSWNV if kind == "cswnv" else S.DSWNV)(**kw)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in g["sd"].items()}, strict=True)
    m = m.cuda().eval()
    rng = np.random.default_rng(3)
    feats = {"/x/u%d.h5" % i: rng.standard_normal((int(t), kw["n_aux"])).astype(np.float32) for i, t in enumerate([3, 5, 2, 5, 4, 7, 6])}
    out = D.gpu_decode(m, list(feats), str(tmp_path), 22050, batch_size=3, reader=lambda f, p: feats[f], seed=11)
    assert sorted(out) == sorted("u%d" % i for i in range(7))
    U, seg = kw["upsampling_factor"], g["cfg"].eff_seg
    # the same batches, straight through the model
    done = 0
    for ids, (bx, bh, ns) in D.decode_generator(list(feats), batch_size=3, seg=seg, upsampling_factor=U, kind=kind,
                                                 reader=lambda f, p: feats[f]):
        direct = m.batch_fast_generate(bx, bh, ns, seed=11 + done)
        done += len(ids)
        for fid, smp, n in zip(ids, direct, ns):
            wav = np.clip(smp, -1, 1) if kind == "cswnv" else S.decode_mu_law(smp, 256)
            assert len(out[fid]) == (n // seg) * seg and np.array_equal(out[fid], wav)
            with wave.open(str(tmp_path / (fid + ".wav")), "rb") as w:
                assert np.array_equal(np.frombuffer(w.readframes(w.getnframes()), np.int16), D.pcm16(wav))
