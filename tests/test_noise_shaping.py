# -*- coding: utf-8 -*-
"""Noise-shaping restore filter (noise_shaping.py:52-87,146-181, feature_extract.py:57-76).

CPU tests pin the oracle: the FIR design against scipy itself, the restated SPTK MLSA filter against the closed-form
transfer function it approximates (pysptk is absent here: see oracle/noise_shaping_oracle.py, "parity unpinned").
GPU tests hold the CUDA path (through the C ABI) to the oracle in float64.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import noise_shaping_oracle as NO  # noqa: E402


def _mean_feat(dim_total=55, seed=3):
    """A mean feature vector shaped like /mean_org_lf0 (uv, lf0, coded ap, mel-cepstrum with a decaying envelope)."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(dim_total)
    v[5:] = 0.6 * rng.standard_normal(dim_total - 5) / (1.0 + 0.25 * np.arange(dim_total - 5))
    v[5] = -4.0
    return v


@pytest.mark.parametrize("fs", [16000, 22050, 24000])
def test_fir_design_matches_scipy(fs):
    firwin = pytest.importorskip("scipy.signal").firwin
    import shallow_wavenet_b200.noise_shaping as NS
    ref = firwin(255, 70 / (fs // 2), pass_zero=False)
    assert np.abs(NO.firwin_highpass(255, 70 / (fs // 2)) - ref).max() < 1e-15
    assert np.abs(NS.firwin_highpass(255, 70 / (fs // 2)) - ref).max() < 1e-15


def test_low_cut_oracle_matches_scipy_lfilter():
    sig = pytest.importorskip("scipy.signal")
    x = np.random.default_rng(0).standard_normal(3000)
    ref = sig.lfilter(sig.firwin(255, 70 / (22050 // 2), pass_zero=False), 1, x)
    assert np.abs(NO.low_cut_filter(x, 22050) - ref).max() < 1e-13


@pytest.mark.parametrize("alpha,pd", [(0.0, 4), (0.455, 4), (0.466, 4), (0.466, 5)])
def test_mlsa_oracle_realises_the_mel_cepstral_transfer_function(alpha, pd):
    """exp(b0) * MLSA(b) approximates H(z) = exp(sum_m c(m) z~^-m): log|H| of the restated filter's impulse response
    equals sum_m c(m) cos(m w~) up to the Pade approximation error (1e-4 class for |log H| < 1)."""
    rng = np.random.default_rng(1)
    order, N = 24, 4096
    mc = np.zeros(order + 1)
    mc[1:] = 0.15 * rng.standard_normal(order) / (1 + 0.3 * np.arange(order))
    mc[0] = 0.05
    x = np.zeros(N); x[0] = 1.0
    T = NO.harvest_num_frames(N, 22050)
    y = NO.synthesize(x, np.tile(NO.mc2b(mc, alpha), [T, 1]), 110, alpha, pd)
    H = np.fft.rfft(y[:3600], 8192)
    assert np.abs(np.log(np.abs(H)) - NO.mlsa_log_magnitude(mc, alpha, 8192)).max() < 5e-4
    assert np.abs(y[3000:3600]).max() < 1e-30                     # the response has died out inside the window


def test_inverse_filter_undoes_the_shaping():
    """--inv true then --inv false (stages 0 and 6 of the recipe) is the identity up to the Pade error."""
    mean = _mean_feat()
    x = 0.1 * np.random.default_rng(2).standard_normal(2400)
    coef = mean[5:] * 0.5; coef[0] = 0.0
    T = NO.harvest_num_frames(len(x), 24000)
    y = NO.synthesis_diff(x, np.tile(np.concatenate([[0.0], -coef[1:]]), [T, 1]), 0.466, 24000)
    z = NO.synthesis_diff(y, np.tile(coef, [T, 1]), 0.466, 24000)
    n = (T - 1) * 120 if (T - 1) * 120 < len(x) else ((len(x) - 1) // 120) * 120
    assert np.abs(z[:n] - x[:n]).max() < 2e-3 * np.abs(x).max()


def test_trailing_partial_frame_and_frame_count():
    x = np.ones(1000)
    b = np.tile(NO.mc2b(np.array([0.0, 0.1, -0.05]), 0.4), [3, 1])        # 3 coefficient frames only
    y = NO.synthesize(x, b, 100, 0.4)
    assert np.all(y[300:] == 0.0) and np.all(y[:300] != 0.0)
    assert NO.harvest_num_frames(24000, 24000) == 201


# ----------------------------------------------------------------------------- GPU: CUDA path vs the oracle
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("fs,alpha", [(24000, 0.466), (22050, 0.455), (16000, 0.41)])
def test_gpu_noise_shaping_matches_oracle(fs, alpha):
    """The whole per-file recipe (mean mel-cepstrum * mag, MLSA restore filter, 70 Hz low cut) for a ragged batch in one
    launch, float64, against the oracle: 1e-11 absolute on signals of amplitude 0.3 (differences: exp()'s last bit)."""
    import shallow_wavenet_b200.noise_shaping as NS
    rng = np.random.default_rng(4)
    mean = _mean_feat(seed=int(fs))
    wavs = [np.clip(0.3 * rng.standard_normal(n), -1, 1) for n in (1500, 977, 2113)]
    outs = NS.noise_shaping(wavs, mean, fs=fs, mcep_alpha=alpha)
    assert len(outs) == 3
    for x, y in zip(wavs, outs):
        ref = NO.noise_shaping(x, mean, fs, 5, alpha)
        assert y.dtype == np.float64 and y.shape == x.shape
        assert np.abs(y - ref).max() < 1e-11, float(np.abs(y - ref).max())
    inv = NS.noise_shaping(wavs[:1], mean, fs=fs, mcep_alpha=alpha, inv=True)[0]
    assert np.abs(inv - NO.noise_shaping(wavs[0], mean, fs, 5, alpha, inv=True)).max() < 1e-11


@gpu
def test_gpu_synthesis_diff_time_varying_coefficients_and_pade5():
    """synthesis_diff with a different mel-cepstrum per frame (linear interpolation inside the frame) against the oracle."""
    import shallow_wavenet_b200.noise_shaping as NS
    rng = np.random.default_rng(5)
    fs, hop, T, order = 22050, 110, 9, 30
    x = 0.2 * rng.standard_normal(T * hop + 37)
    mc = 0.2 * rng.standard_normal((T, order + 1)) / (1 + 0.2 * np.arange(order + 1))
    y = NS.synthesis_diff(x, mc, alpha=0.455, fs=fs)
    ref = NO.synthesis_diff(x, mc, 0.455, fs)
    assert np.abs(y - ref).max() < 1e-11
    assert np.all(y[T * hop:] == 0.0)                             # the trailing partial frame is not synthesised
    y5 = NS._run([x], mc, None, 0.455, hop, None, pade=5)[0]
    assert np.abs(y5 - NO.synthesis_diff(x, mc, 0.455, fs, pd=5)).max() < 1e-11


@gpu
def test_gpu_low_cut_filter_matches_scipy():
    sig = pytest.importorskip("scipy.signal")
    import shallow_wavenet_b200.noise_shaping as NS
    x = np.random.default_rng(6).standard_normal(5000)
    ref = sig.lfilter(sig.firwin(255, 70 / (24000 // 2), pass_zero=False), 1, x)
    assert np.abs(NS.low_cut_filter(x, 24000) - ref).max() < 1e-12


@gpu
def test_gpu_noise_shaping_main_writes_pcm16(tmp_path):
    """main() with the script's flags: wav files in, filtered PCM16 wav files out (stdlib wav fallback when soundfile is absent)."""
    import wave
    import shallow_wavenet_b200.noise_shaping as NS
    from shallow_wavenet_b200.decode import write_wav
    rng = np.random.default_rng(7)
    src = tmp_path / "in"; dst = tmp_path / "out"; src.mkdir()
    mean = _mean_feat()
    np.save(str(tmp_path / "stats.npy"), mean)
    wavs = {}
    for name, n in (("a.wav", 2400), ("b.wav", 3000)):
        w = np.clip(0.2 * rng.standard_normal(n), -1, 1)
        write_wav(str(src / name), w, 24000)
        wavs[name] = NS._read_wav(str(src / name))[0]
    assert NS.main(["--waveforms", str(src), "--stats", str(tmp_path / "stats.npy"), "--writedir", str(dst), "--fs", "24000",
                    "--mcep_alpha", "0.466", "--inv", "false"]) == 0
    for name, x in wavs.items():
        ref = NO.noise_shaping(x, mean, 24000, 5, 0.466)
        with wave.open(str(dst / name), "rb") as f:
            got = np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").astype(np.int64)
        assert np.abs(got - np.rint(np.clip(ref, -1, 1) * 32767.0)).max() <= 1
