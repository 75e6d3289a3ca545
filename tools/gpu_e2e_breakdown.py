"""Where the end-to-end call spends its time beyond the kernel (cfg2): phases of DSWNV.batch_fast_generate."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from shallow_wavenet_b200 import synth
kind, fs, over, B, frames, desc = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
kw = bench.model_kwargs(kind, fs, over)
model = synth.make_model(kind, **kw).cuda()
aux_host = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).pin_memory()
n = frames * kw["upsampling_factor"]
seed_audio = torch.full((B, 1), 128, dtype=torch.long) if kind == "dswnv" else torch.zeros(B, kw.get("seg", 1))
eng = model.engine()
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    outs = model.batch_fast_generate(seed_audio, aux_host, [n] * B, seed=5 + it)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    i = eng.get_info()
    print("e2e call %.2f ms | kernel %.2f ms cond %.3f ms" % ((t1 - t0) * 1e3, i["last_gen_ms"], i["last_cond_ms"]))
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e = model.engine(); t_eng = time.perf_counter()
    aux = aux_host.to("cuda", torch.float32).contiguous(); torch.cuda.synchronize(); t_h2d = time.perf_counter()
    e.condition(aux); torch.cuda.synchronize(); t_cond = time.perf_counter()
    out, _ = e.generate(n, seed=9 + it); torch.cuda.synchronize(); t_gen = time.perf_counter()
    s = out.cpu().numpy(); t_d2h = time.perf_counter()
    i = e.get_info()
    print("engine() %.2f | h2d %.2f | condition %.2f | generate+sync %.2f (kernel %.2f) | d2h %.2f ms" % (
        (t_eng - t0) * 1e3, (t_h2d - t_eng) * 1e3, (t_cond - t_h2d) * 1e3, (t_gen - t_cond) * 1e3, i["last_gen_ms"], (t_d2h - t_gen) * 1e3))
