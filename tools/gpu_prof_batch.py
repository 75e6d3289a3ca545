"""Stage timeline of the pipelined batch kernel (variant 3) from in-kernel %globaltimer stamps.
usage: gpu_prof_batch.py [cswnv|dswnv] [B] [tile]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shallow_wavenet_b200 import synth

kind = sys.argv[1] if len(sys.argv) > 1 else "cswnv"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 8
fs = 22050 if kind == "cswnv" else 24000
kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
model = synth.make_model(kind, **kw).cuda()
frames = 10
aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
eng = model.engine()
n = frames * kw["upsampling_factor"]
eng.condition(aux)
eng.generate(n, seed=0, batch_tile=tile, kernel_sel=int(os.environ.get('ENGINE', '3')))
info = eng.info
M, grid, G, NHEAD, L = info["n_slots"], info["grid"], info["ctas_per_layer"], info["head_ctas"], 6
start, count = (info["receptive_field"] // info["eff_seg"]) + 60, 3
prof = torch.zeros((count, M * grid // 2, 32), dtype=torch.int64, device="cuda")
eng.condition(aux)
eng.generate(n, seed=0, batch_tile=tile, kernel_sel=int(os.environ.get('ENGINE', '3')), prof=prof, prof_start=start)
p = prof.cpu().numpy().reshape(count, M, grid, 16).astype(np.float64)
print("%s B=%d tile=%d slots=%d grid=%d G=%d head CTAs=%d kernel %.2f ms" % (kind, B, tile, M, grid, G, NHEAD, eng.info["last_gen_ms"]))
stamps = p[..., :8]
t00 = stamps[stamps > 0].min()
it = 1
nshow = int(os.environ.get("NSLOTS", "3"))
print("iteration %d: per slot, first CTA of each cell group; us relative to the first stamp; g[start done] / pub / P3 relative to in-ready" % (start + it))
for m in range(min(M, nshow)):
    row = []
    for l in range(L):
        e = p[it, m, l * G]
        row.append("L%d in %7.2f g[%5.2f %5.2f] pub %5.2f P3 %5.2f" % (
            l, (e[0] - t00) / 1e3, (e[1] - e[0]) / 1e3, (e[2] - e[0]) / 1e3, (e[5] - e[0]) / 1e3, (e[6] - e[0]) / 1e3))
    print("slot %2d | " % m + " | ".join(row))
    e = p[it, m, 0]
    if e[8] > 0 and e[7] > 0:
        print("        | cell-0 builder (warp 0): samples ready 0, taps staged %.2f, tile computed %.2f, fenced+arrived %.2f, MMA start %.2f" % (
            (e[7] - e[0]) / 1e3, (e[8] - e[0]) / 1e3, (e[9] - e[0]) / 1e3, (e[1] - e[0]) / 1e3))
    hd = []
    for c in range(L * G, grid):
        e = p[it, m, c]
        if e[0] == 0:
            continue
        if e[7] == 0:      # dense stage CTA (only the tail stamps event 7)
            hd.append("D%d in %7.2f g[%5.2f %5.2f] pub %5.2f" % (c - L * G, (e[0] - t00) / 1e3, (e[1] - e[0]) / 1e3, (e[2] - e[0]) / 1e3, (e[5] - e[0]) / 1e3))
        else:
            hd.append("tail in %7.2f x2-loaded %5.2f logits %5.2f sampled %5.2f input: computed %5.2f synced %5.2f signalled %5.2f" % (
                (e[0] - t00) / 1e3, (e[7] - e[0]) / 1e3, (e[1] - e[0]) / 1e3, (e[3] - e[0]) / 1e3, (e[5] - e[0]) / 1e3, (e[6] - e[0]) / 1e3, (e[4] - e[0]) / 1e3))
    print("        | " + " | ".join(hd))
lay = p[:, :, :L * G, :]
ok = lay[..., 6] > 0
def avg(a, b):
    d = (lay[..., b] - lay[..., a])[ok]
    return d.mean() / 1e3, np.percentile(d, 95) / 1e3
for a, b, nm in ((0, 1, "in-ready -> gemm-start (TMA)"), (1, 2, "gemm"), (2, 5, "gemm-done -> published"), (5, 6, "P3 (ring scatter)"), (0, 6, "whole pass")):
    print("%-32s mean %6.2f us  p95 %6.2f us" % (nm, *avg(a, b)))
per = np.diff(np.sort(lay[:, :, 0, 0].reshape(-1)))
print("pass period of CTA 0: mean %.2f us ; iteration period %.2f us" % (per.mean() / 1e3, per.mean() * M / 1e3))
