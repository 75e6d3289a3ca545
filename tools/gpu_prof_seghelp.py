"""Helper CTAs of the seg cluster kernel with B utterances side by side: per layer, when the inputs of cluster q were in and
when its sums were handed over (stamps 2q / 2q+1; library built with -DSWN_STAMPS=1, tools/seg_build.sh)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from shallow_wavenet_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
kind, fs, over, _, frames, desc = bench.WORKLOADS["cfg1"]
kw = bench.model_kwargs(kind, fs, over)
model = synth.make_model(kind, **kw).cuda()
frames = 40
aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
eng = model.engine()
n = frames * kw["upsampling_factor"]
eng.condition(aux)
eng.generate(n, seed=1, kernel_sel=2)
i0 = eng.info
print("== B=%d unprofiled: %.2f ms, grid %d, side by side %d" % (B, i0["last_gen_ms"], i0["grid"], i0["batch_tile"]))
grid, cnt = i0["grid"], 64
prof = torch.zeros((cnt, grid, 32), dtype=torch.int64, device="cuda")
eng.condition(aux)
eng.generate(n, seed=1, kernel_sel=2, prof=prof, prof_start=800)
print("== with stamps: %.2f ms" % eng.info["last_gen_ms"])
p = prof.cpu().numpy().astype(np.float64)[8:-1]
ncl = i0["batch_tile"]
core0 = p[:, 0, 0]
print("   core step period %.0f ns" % np.diff(core0).mean())
for c in range(ncl):
    print("   core cluster %d: iter start relative to cluster 0: %+.0f ns; L1..L5 input ready %s" % (
        c, (p[:, 16 * c, 0] - core0).mean(), " ".join("%.0f" % (p[:, 16 * c, e] - p[:, 16 * c, 0]).mean() for e in range(1, 7))))
h0 = 16 * ncl
rows = []
for b in range(h0, grid):
    if p[:, b, 0].max() <= 0:
        continue
    t_in = [p[:, b, 2 * q] for q in range(ncl)]
    t_out = [p[:, b, 2 * q + 1] for q in range(ncl)]
    rows.append((b, [float((t_out[q] - t_in[q]).mean()) for q in range(ncl)], float((t_out[ncl - 1] - t_in[0]).mean()),
                 [float((t_in[q] - core0).mean()) for q in range(ncl)]))
for b, comp, span, tin in rows:
    print("   helper CTA %3d: compute per cluster %s ns; first in -> last out %.0f ns; inputs in at %s ns after cluster 0's step start" % (
        b, " ".join("%.0f" % c for c in comp), span, " ".join("%.0f" % t for t in tin)))
