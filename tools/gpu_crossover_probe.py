"""Which persistent kernel wins at small batches?  (v1 layer groups / v2 cluster / v3 slot ring)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shallow_wavenet_b200 import synth

frames = int(os.environ.get("SWN_PROBE_FRAMES", 20))
for kind, fs, over in (("cswnv", 22050, dict(seg=5, lpc=4)), ("dswnv", 16000, dict())):
    kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
    kw.update(over)
    model = synth.make_model(kind, **kw).cuda()
    eng = model.engine()
    n = frames * kw["upsampling_factor"]
    for B in (1, 2, 3, 4, 6, 8, 12, 16, 32):
        aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
        row = []
        for sel in (1, 2, 3):
            try:
                for rep in range(2):
                    eng.condition(aux)
                    eng.generate(n, seed=rep, kernel_sel=sel)
                info = eng.info
                row.append("sel %d: var %d tile %2d slots %2d %8.2f ms %9.0f samples/s" % (
                    sel, info["kernel_variant"], info["batch_tile"], info["n_slots"], info["last_gen_ms"],
                    B * (n + info["receptive_field"]) / info["last_gen_ms"] * 1e3))
            except Exception as e:
                row.append("sel %d: n/a" % sel)
        print("%s B=%2d | " % (kind, B) + " | ".join(row), flush=True)
