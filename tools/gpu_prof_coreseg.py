"""Timeline of the seg > 1 cluster kernel (core CTAs) from %globaltimer stamps (library built with -DSWN_STAMPS=1)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from shallow_wavenet_b200 import synth

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
kind, fs, over, B, frames, desc = bench.WORKLOADS[wl]
kw = bench.model_kwargs(kind, fs, over)
model = synth.make_model(kind, **kw).cuda()
aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, 1)).cuda()
eng = model.engine()
n = frames * kw["upsampling_factor"]
seg = kw.get("seg", 1)
eng.condition(aux)
eng.generate(n, seed=1, kernel_sel=2)
i0 = eng.info
steps = n // seg + (i0["receptive_field"] + seg) // seg
print("== %s kernel_variant=%d unprofiled: kernel %.2f ms, %.3f us/step (%d steps of %d samples), %.0f samples/s" % (
    wl, i0["kernel_variant"], i0["last_gen_ms"], i0["last_gen_ms"] * 1e3 / steps, steps, seg, n / i0["last_gen_ms"] * 1e3))
grid = i0["grid"]
cnt = 64
prof = torch.zeros((cnt, grid, 32), dtype=torch.int64, device="cuda")
eng.condition(aux)
eng.generate(n, seed=1, kernel_sel=2, prof=prof, prof_start=1500)
info = eng.info
print("== with stamps: kernel %.2f ms, %.3f us/step" % (info["last_gen_ms"], info["last_gen_ms"] * 1e3 / steps))
pall = prof.cpu().numpy().astype(np.float64)[8:-1, :16]
base = pall[:, :, 0].min(axis=1, keepdims=True)[:, :, None]
rel = pall - base
print("   step period %.0f ns" % np.diff(pall[:, 0, 0]).mean())
names = {0: "iter start", 1: "L0 input ready", 2: "L1 input ready", 3: "L2 input ready", 4: "L3 input ready", 5: "L4 input ready",
         6: "L5 input ready", 7: "layer section left (warp 0)", 9: "out_1 row published (warp 0)", 16: "out_2 row done (warp 0)",
         10: "sampled + CTA barrier", 11: "next inputs published", 12: "step end (warp 0)"}
for e in (0, 1, 2, 3, 4, 5, 6, 7, 9, 16, 10, 11, 12):
    if pall[:, :, e].max() > 0:
        print("   %-32s mean %6.0f ns  (min CTA %6.0f, max CTA %6.0f)" % (names[e], rel[:, :, e].mean(), rel[:, :, e].mean(0).min(), rel[:, :, e].mean(0).max()))
for e in range(17, 32):
    if pall[:, :, e].max() > 0:
        print("   stamp %2d                         mean %6.0f ns  (min CTA %6.0f, max CTA %6.0f)" % (e, rel[:, :, e].mean(), rel[:, :, e].mean(0).min(), rel[:, :, e].mean(0).max()))
