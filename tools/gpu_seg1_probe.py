"""CSWNV seg=1 (the recipe's default: model=laplace, seg=1, lpc=4, H=192) at small batches: cluster kernel (launches of two) against
the batch kernel and the automatic choice."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shallow_wavenet_b200 import synth

frames = int(os.environ.get("SWN_PROBE_FRAMES", 20))
kw = synth.cswnv_kwargs(22050)
kw.update(dict(seg=1, lpc=4))
model = synth.make_model("cswnv", **kw).cuda()
eng = model.engine()
n = frames * kw["upsampling_factor"]
for B in (1, 2, 3, 4, 5, 6, 7, 8, 12):
    aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
    row = []
    for sel in (0, 2, 3):
        try:
            for rep in range(2):
                eng.condition(aux)
                out = eng.generate(n, seed=7, kernel_sel=sel)
            out = out[0] if isinstance(out, tuple) else out
            info = eng.info
            row.append("sel %d: var %d tile %2d slots %2d grid %3d %8.2f ms %9.0f samples/s  sum|x| %.5f" % (
                sel, info["kernel_variant"], info["batch_tile"], info["n_slots"], info["grid"], info["last_gen_ms"],
                B * (n + info["receptive_field"]) / info["last_gen_ms"] * 1e3, float(torch.as_tensor(out).abs().double().sum())))
        except Exception as e:
            row.append("sel %d: n/a (%s)" % (sel, str(e)[:60]))
    print("cswnv seg1 B=%2d | " % B + " | ".join(row), flush=True)
