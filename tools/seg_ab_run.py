"""Dev aid: time the seg cluster kernel of every library shallow-wavenet_b200/lib/ab/libswn_seg_*.so on one workload (one subprocess each)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r)
import torch, bench
from shallow_wavenet_b200 import synth
wl = sys.argv[1]
kind, fs, over, B, frames, desc = bench.WORKLOADS[wl]
kw = bench.model_kwargs(kind, fs, over)
model = synth.make_model(kind, **kw).cuda()
aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, 1)).cuda()
eng = model.engine()
n = frames * kw["upsampling_factor"]
seg = kw.get("seg", 1)
ts = []
for r in range(4):
    eng.condition(aux); out = eng.generate(n, seed=1, kernel_sel=2)[0]; i = eng.info
    ts.append(i["last_gen_ms"] * 1e3 / (n // seg + (i["receptive_field"] + seg) // seg))
print("%%-36s variant %%d  us/step %%s  checksum %%d" %% (os.path.basename(os.environ.get("SWN_B200_LIB", "default")), i["kernel_variant"],
      " ".join("%%.3f" %% t for t in ts), int(out.double().abs().sum().item() * 1e3)))
''' % ROOT
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
libs = sorted(glob.glob(os.path.join(ROOT, "shallow-wavenet_b200", "lib", "ab", "libswn_seg_*.so")))
for lib in [None] + [l for l in libs if "stamps" not in l]:
    env = dict(os.environ)
    if lib: env["SWN_B200_LIB"] = lib
    subprocess.run([sys.executable, "-c", CHILD, wl], env=env)
