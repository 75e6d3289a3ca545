"""Run-to-run reproducibility of the one-sample-per-step cluster kernel (CSWNV seg 1, H=192, B=2, teacher-forced heads): three
identical runs, max |difference| of the heads against the first.  Used with the late-helper debug builds of tools/core3_build.sh."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shallow_wavenet_b200 import synth
kw = synth.cswnv_kwargs(22050); kw.update(dict(seg=1, lpc=4))
model = synth.make_model("cswnv", **kw).cuda()
eng = model.engine()
frames, B = 6, int(os.environ.get("SWN_PROBE_B", 2))
n = frames * kw["upsampling_factor"]
aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
forced = (0.3 * np.random.default_rng(0).standard_normal((B, n))).astype(np.float32)
res = []
for rep in range(4):
    eng.condition(aux)
    out, heads = eng.generate(n, seed=7, kernel_sel=2, forced=forced, return_heads=True)
    res.append(heads.cpu().numpy().copy())
d = [float(np.abs(r - res[0]).max()) for r in res[1:]]
first = [int(np.nonzero(np.abs(r - res[0]).reshape(B, n, -1).max(-1).max(0) > 0)[0][:1].sum()) for r in res[1:]]
last = [int(np.nonzero(np.abs(r - res[0]).reshape(B, n, -1).max(-1).max(0) > 0)[0][-1:].sum()) for r in res[1:]]
ref_path = os.environ.get("SWN_PROBE_REF")
if ref_path and os.path.exists(ref_path):
    ref = np.load(ref_path)
    dd = np.abs(res[0] - ref).reshape(B, n, -1).max(-1).max(0)
    nz = np.nonzero(dd > 0)[0]
    print("   against %s: max %.3e, differing steps %d (first %s, last %s)" % (ref_path, dd.max(), len(nz), nz[:1], nz[-1:]))
elif ref_path:
    np.save(ref_path, res[0])
print("%s: %.2f ms, max head diff vs run 0: %s, first/last differing step: %s %s" % (
    os.path.basename(os.environ.get("SWN_B200_LIB", "default")), eng.info["last_gen_ms"], d, first, last))
