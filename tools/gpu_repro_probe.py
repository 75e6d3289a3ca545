"""Cluster kernels with artificially late helper CTAs (test hook swn_gen_args::debug_helper_delay_ns): teacher-forced heads must be
bit-identical to the undelayed run.  This is how the missing reconvergence after the operand commit of swn_core2.cuh was found
(profiles/r03_core2_late_helper_repro.txt); tests/test_gpu_parity.py keeps a short version of it."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shallow_wavenet_b200 import synth

frames = 6
for kind, over, Bs in (("cswnv", dict(seg=1, lpc=4), (1, 2)), ("dswnv", dict(), (1, 2)), ("cswnv", dict(seg=5, lpc=4), (1, 3))):
    kw = synth.cswnv_kwargs(22050) if kind == "cswnv" else synth.dswnv_kwargs(16000)
    kw.update(over)
    model = synth.make_model(kind, **kw).cuda()
    eng = model.engine()
    n = frames * kw["upsampling_factor"]
    for B in Bs:
        aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
        rng = np.random.default_rng(B)
        forced = (0.3 * rng.standard_normal((B, n))).astype(np.float32) if kind == "cswnv" else rng.integers(0, 256, (B, n))
        ref = None
        for delay in (0, 4000, 8000, 20000):
            eng.condition(aux)
            out, heads = eng.generate(n, seed=7, kernel_sel=2, forced=forced, return_heads=True, helper_delay_ns=delay)
            h = heads.cpu().numpy()
            ref = h if ref is None else ref
            print("%s %s B=%d helper delay %5d ns: %7.2f ms, max |head - undelayed| %.3e" % (
                kind, over, B, delay, eng.info["last_gen_ms"], float(np.abs(h - ref).max())), flush=True)
