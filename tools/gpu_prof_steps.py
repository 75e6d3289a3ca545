"""Per-stage timeline of the persistent kernel from in-kernel %globaltimer stamps (GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from shallow_wavenet_b200 import synth

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2s"
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2]
kind, fs, over, B, frames, desc = bench.WORKLOADS[wl]
kw = bench.model_kwargs(kind, fs, over)
model = synth.make_model(kind, **kw).cuda()
aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
eng = model.engine()
n = frames * kw["upsampling_factor"]
L = kw["dilation_depth"] * kw["dilation_repeat"]
for mode in modes:
    eng.condition(aux)
    eng.generate(n, seed=1, poll_mode=mode)           # warm
    grid = eng.info["grid"]
    cnt = 64
    prof = torch.zeros((cnt, grid, 32), dtype=torch.int64, device="cuda")
    eng.condition(aux)
    out, _ = eng.generate(n, seed=1, poll_mode=mode, prof=prof, prof_start=3000)
    info = eng.info
    seg = info["eff_seg"]
    steps = n // seg + info["receptive_field"] // seg + 1
    print("== %s poll_mode=%d kernel %.2f ms, %.3f us/iteration, %.0f samples/s, grid %d G=%d" % (
        wl, mode, info["last_gen_ms"], info["last_gen_ms"] * 1e3 / steps, B * n / info["last_gen_ms"] * 1e3, grid, info["ctas_per_layer"]))
    p = prof.cpu().numpy().astype(np.float64)          # ns
    G = info["ctas_per_layer"]
    rows = []
    for it in range(8, cnt - 1):
        base = p[it, L * G:, 0].max()                  # head: last layer output ready
        t = []
        for l in range(L):
            blk = p[it, l * G:(l + 1) * G]
            t.append((blk[:, 0].mean() - base, blk[:, 0].max() - base, blk[:, 1].mean() - base, blk[:, 1].max() - base, blk[:, 2].max() - base))
        hb = p[it, L * G:]
        rows.append((t, (hb[:, 0].mean() - base, hb[:, 1].max() - base, hb[:, 2].max() - base, hb[:, 3].max() - base),
                     p[it + 1, L * G:, 0].max() - base))
    tm = np.mean([[x for tt in r[0] for x in tt] for r in rows], 0).reshape(L, 5)
    hm = np.mean([r[1] for r in rows], 0)
    per = np.mean([r[2] for r in rows])
    print("   iteration period %.0f ns" % per)
    print("   head: in_ready(mean) %.0f  A_done %.0f  B_done %.0f  next_input_published %.0f" % tuple(hm))
    for l in range(L):
        print("   layer %d: in_ready mean %.0f max %.0f | published mean %.0f max %.0f | passB done max %.0f  (ns rel. to head-ready)" % ((l,) + tuple(tm[l])))
