"""Throughput probe of batched workloads (short utterances) on the GPU box."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shallow_wavenet_b200 import synth

cases = [("cswnv", 22050, dict(seg=5, lpc=4), B, 20) for B in (1, 2, 8, 64)] + \
        [("dswnv", 16000, dict(), B, 20) for B in (1, 2, 4, 8, 32, 128)]
for kind, fs, over, B, frames in cases:
    kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
    kw.update(over)
    model = synth.make_model(kind, **kw).cuda()
    aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
    eng = model.engine()
    n = frames * kw["upsampling_factor"]
    try:
        for rep in range(2):
            eng.condition(aux)
            out, _ = eng.generate(n, seed=rep)
        info = eng.info
        seg = info["eff_seg"]
        steps = n // seg + info["receptive_field"] // seg + 1
        print("%s B=%3d: variant %d tile %d slots %3d grid %d smem %6d | kernel %8.2f ms cond %6.2f ms | %7.2f us/iteration | %9.0f samples/s total, %6.2fx RT per utterance" % (
            kind, B, info["kernel_variant"], info["batch_tile"], info["n_slots"], info["grid"], info["smem_bytes"], info["last_gen_ms"], info["last_cond_ms"],
            info["last_gen_ms"] * 1e3 / steps, B * n / info["last_gen_ms"] * 1e3, n / info["last_gen_ms"] * 1e3 / fs))
    except Exception as e:
        print("%s B=%d FAILED: %s" % (kind, B, e))
    del model, eng
    torch.cuda.empty_cache()
