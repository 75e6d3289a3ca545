"""Dev aid: time the un-profiled batch-1 kernel of every library under shallow-wavenet_b200/lib/ab (one subprocess each)."""
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
aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
eng = model.engine()
n = frames * kw["upsampling_factor"]
ts = []
for r in range(4):
    eng.condition(aux); out = eng.generate(n, seed=1)[0]; i = eng.info
    ts.append(i["last_gen_ms"] * 1e3 / (n + i["receptive_field"] + 1))
print("%%-28s variant %%d  us/step %%s  checksum %%d" %% (os.path.basename(os.environ.get("SWN_B200_LIB", "default")), i["kernel_variant"],
      " ".join("%%.3f" %% t for t in ts), int(out.sum().item()) if out.dtype == torch.int64 else int(out.double().abs().sum().item() * 1e3)))
''' % ROOT
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2s"
objdir = os.path.join(ROOT, "shallow-wavenet_b200", "lib", "obj")
common = [o for o in sorted(glob.glob(os.path.join(objdir, "*.o"))) if not o.endswith("swn_core_nch4.o")]
libs = []
for o in sorted(glob.glob(os.path.join(ROOT, "shallow-wavenet_b200", "lib", "ab", "core4_*.o"))):
    lib = "/tmp/libswn_%s.so" % os.path.basename(o)[6:-2]
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + common + [o, "-cudart", "static"], check=True)
    libs.append(lib)
for lib in [None] + libs:
    env = dict(os.environ)
    if lib: env["SWN_B200_LIB"] = lib
    subprocess.run([sys.executable, "-c", CHILD, wl], env=env)
