# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE (never imported by the product path): CPU restatement of the noise-shaping restore filter,
SURVEY.md 8f rank 4 - reference ``src/bin/noise_shaping.py:52-87`` (``synthesis_diff``) and ``:165-181`` (the per-file
recipe), ``src/bin/feature_extract.py:57-76`` (``low_cut_filter``).

PARITY STATUS.  The reference delegates the arithmetic to third-party code that is ABSENT from /root/reference and from
this image: **pysptk** (``tools/requirements.txt:7``, unpinned; the PyPI releases wrap **SPTK 3.11**) for ``mc2b``,
``MLSADF`` and ``Synthesizer``, **scipy.signal** for the low-cut FIR, **pyworld.harvest** for the frame count.
  * MLSA part - **parity unpinned**: this file restates SPTK 3.11's published algorithm (``mc2b.c``, ``mlsadf.c``:
    ``mlsadf1`` / ``mlsadf2`` / ``mlsafir`` with the Pade table of order 4 / 5) and pysptk's ``Synthesizer`` (per-sample linear
    interpolation of the filter coefficients inside a frame, input scaled by ``exp(b[0])``, trailing partial frame not
    synthesised).  There is no pysptk here to generate golden vectors from.  What the tests do instead: the impulse response
    of the restated filter is held to the closed form it approximates, ``log|H(w)| = sum_m c(m) cos(m w~)`` with the all-pass
    warped frequency ``w~`` (a wrong Pade coefficient, sign, tap alignment or ``mc2b`` recursion fails that test).
  * low-cut part - **pinned**: ``firwin_highpass`` restates scipy's ``firwin(255, cutoff, pass_zero=False)`` and is checked
    against scipy itself (installed in this image) to 1e-15; ``lfilter(fil, 1, x)`` is a plain causal FIR.
  * frame count: ``pyworld.harvest`` returns ``int(1000 * len / fs / frame_period) + 1`` frames (its documented length
    formula); only the count is used (``noise_shaping.py:160-161``).
"""
import numpy as np

# SPTK 3.11 mlsadf.c: Pade approximants of exp(x), orders 1..5 stored back to back; row pd starts at pd*(pd+1)/2
PADE = np.array([1.0,
                 1.0, 0.0,
                 1.0, 0.0, 0.0,
                 1.0, 0.0, 0.0, 0.0,
                 1.0, 0.4999273, 0.1067005, 0.01170221, 0.0005656279,
                 1.0, 0.4999391, 0.1107098, 0.01369984, 0.0009564853, 0.00003041721], dtype=np.float64)


def mc2b(mc, alpha):
    """SPTK mc2b.c: b(m) = c(m); b(i) = c(i) - alpha * b(i+1)."""
    mc = np.asarray(mc, dtype=np.float64)
    b = np.empty_like(mc)
    m = len(mc) - 1
    b[m] = mc[m]
    for i in range(m - 1, -1, -1):
        b[i] = mc[i] - alpha * b[i + 1]
    return b


def harvest_num_frames(n_samples, fs, frame_period=5.0):
    """Length of pyworld.harvest's f0 track (noise_shaping.py:160-161 uses nothing else of it)."""
    return int(1000.0 * n_samples / fs / frame_period) + 1


class MLSADF(object):
    """pysptk.synthesis.MLSADF(order, alpha, pd) = SPTK mlsadf() with its delay buffer."""

    def __init__(self, order, alpha, pd=4):
        assert pd in (4, 5)
        self.m, self.a, self.pd = order, float(alpha), pd
        self.d = np.zeros(3 * (pd + 1) + pd * (order + 2), dtype=np.float64)
        self.ppade = PADE[pd * (pd + 1) // 2:]

    def _mlsafir(self, x, b, d):
        m, a = self.m, self.a
        aa = 1.0 - a * a
        y = 0.0
        d[0] = x
        d[1] = aa * d[0] + a * d[1]
        for i in range(2, m + 1):
            d[i] = d[i] + a * (d[i + 1] - d[i - 1])
            y += d[i] * b[i]
        for i in range(m + 1, 1, -1):
            d[i] = d[i - 1]
        return y

    def _mlsadf1(self, x, b, d):
        a, pd = self.a, self.pd
        aa = 1.0 - a * a
        out = 0.0
        pt = d[pd + 1:]
        for i in range(pd, 0, -1):
            d[i] = aa * pt[i - 1] + a * d[i]
            pt[i] = d[i] * b[1]
            v = pt[i] * self.ppade[i]
            x += v if (i & 1) else -v
            out += v
        pt[0] = x
        out += x
        return out

    def _mlsadf2(self, x, b, d):
        m, pd = self.m, self.pd
        out = 0.0
        pt = d[pd * (m + 2):]
        for i in range(pd, 0, -1):
            pt[i] = self._mlsafir(pt[i - 1], b, d[(i - 1) * (m + 2):])
            v = pt[i] * self.ppade[i]
            x += v if (i & 1) else -v
            out += v
        pt[0] = x
        out += x
        return out

    def filt(self, x, b):
        x = self._mlsadf1(x, b, self.d)
        return self._mlsadf2(x, b, self.d[2 * (self.pd + 1):])


def synthesize(x, b, hop, alpha, pd=4):
    """pysptk.synthesis.Synthesizer(MLSADF(order, alpha, pd), hop).synthesis(x, b): per frame the coefficients move
    linearly from the previous frame's to this frame's (the first frame is constant), the input is scaled by
    exp(b[0]); a trailing partial frame is not synthesised (pysptk leaves it uninitialised - zeros here)."""
    x = np.asarray(x, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    order = b.shape[1] - 1
    f = MLSADF(order, alpha, pd)
    y = np.zeros_like(x)
    prev = b[0].copy()
    for i in range(len(b)):
        s, e = i * hop, (i + 1) * hop
        if e >= len(x):
            break
        curr = b[i]
        slope = (curr - prev) / hop
        coef = prev.copy()
        for n in range(s, e):
            y[n] = f.filt(x[n] * np.exp(coef[0]), coef)
            coef += slope
        prev = curr
    return y


def synthesis_diff(x, diffmcep, alpha, fs, shiftms=5.0, pd=4):
    """noise_shaping.py:52-87 with rmcep=None."""
    x = np.asarray(x, dtype=np.float64)
    shiftl = int(fs / 1000 * shiftms)
    b = np.apply_along_axis(mc2b, 1, np.asarray(diffmcep, dtype=np.float64), alpha)
    assert np.isfinite(b).all()
    return synthesize(x, b, shiftl, alpha, pd)


def firwin_highpass(numtaps, cutoff_norm):
    """scipy.signal.firwin(numtaps, cutoff_norm, pass_zero=False) (Hamming window, scale=True): pass band
    [cutoff, Nyquist]; odd numtaps."""
    assert numtaps % 2 == 1
    alpha = 0.5 * (numtaps - 1)
    m = np.arange(0, numtaps) - alpha
    left, right = float(cutoff_norm), 1.0
    h = right * np.sinc(right * m) - left * np.sinc(left * m)
    h = h * np.hamming(numtaps)
    # unit gain at the centre of the first pass band, which ends at Nyquist: scale frequency 1.0
    scale_frequency = 1.0
    c = np.cos(np.pi * m * scale_frequency)
    return h / np.sum(h * c)


def low_cut_filter(x, fs, cutoff=70):
    """feature_extract.py:57-76: firwin(255, cutoff / (fs // 2), pass_zero=False), lfilter(fil, 1, x)."""
    fil = firwin_highpass(255, cutoff / (fs // 2))
    x = np.asarray(x, dtype=np.float64)
    return np.convolve(x, fil)[:len(x)]


def noise_shaping(x, mean_feat, fs, mcep_dim_start, mcep_alpha, mag=0.5, inv=False, shiftms=5.0):
    """The per-file recipe of noise_shaping.py:146-181 between reading and writing the wav file."""
    x = np.asarray(x, dtype=np.float64)
    num_frames = harvest_num_frames(len(x), fs)
    coef = np.array(mean_feat, dtype=np.float64)[mcep_dim_start:] * mag
    coef[0] = 0.0
    if inv:
        coef[1:] = -1.0 * coef[1:]
    coef = np.tile(coef, [num_frames, 1])
    y = synthesis_diff(x, coef, alpha=mcep_alpha, fs=fs, shiftms=shiftms)
    return low_cut_filter(y, fs, cutoff=70)


def mlsa_log_magnitude(mc, alpha, n_fft):
    """The transfer function the MLSA filter approximates: log|H(e^jw)| = sum_m c(m) cos(m w~), w~ = the phase of the
    first-order all-pass (z^-1 - alpha) / (1 - alpha z^-1) at w.  (exp(b[0]) gain included through c(0).)"""
    w = np.arange(n_fft // 2 + 1) * 2.0 * np.pi / n_fft
    wt = w + 2.0 * np.arctan2(alpha * np.sin(w), 1.0 - alpha * np.cos(w))
    mc = np.asarray(mc, dtype=np.float64)
    return sum(mc[k] * np.cos(k * wt) for k in range(len(mc)))
