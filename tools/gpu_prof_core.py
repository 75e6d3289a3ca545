"""Timeline of the v2 cluster kernel (core CTAs) from %globaltimer stamps; also prints throughput."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from shallow_wavenet_b200 import synth

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2s"
kind, fs, over, B, frames, desc = bench.WORKLOADS[wl]
B = int(os.environ.get("SWN_PROF_B", B))          # e.g. SWN_PROF_B=2: the two-column instantiation
kw = bench.model_kwargs(kind, fs, over)
model = synth.make_model(kind, **kw).cuda()
aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
eng = model.engine()
n = frames * kw["upsampling_factor"]
for sel in ((2,) if os.environ.get("SWN_PROF_B") else (2, 1)):
    eng.condition(aux)
    try:
        eng.generate(n, seed=1, kernel_sel=sel)
    except Exception as e:
        print("kernel_sel", sel, "failed:", e); continue
    grid = eng.info["grid"]
    i0 = eng.info
    print("== %s kernel_variant=%d unprofiled: kernel %.2f ms, %.3f us/step" % (
        wl, i0["kernel_variant"], i0["last_gen_ms"], i0["last_gen_ms"] * 1e3 / (n + i0["receptive_field"] + 1)))
    cnt = 64
    prof = torch.zeros((cnt, grid, 32), dtype=torch.int64, device="cuda")
    eng.condition(aux)
    eng.generate(n, seed=1, kernel_sel=sel, prof=prof, prof_start=3000)
    info = eng.info
    steps = n + info["receptive_field"] + 1
    print("== %s kernel_variant=%d kernel %.2f ms, %.3f us/step, %.0f samples/s (RTF %.1f), grid %d smem %d" % (
        wl, info["kernel_variant"], info["last_gen_ms"], info["last_gen_ms"] * 1e3 / steps, B * n / info["last_gen_ms"] * 1e3,
        B * n / info["last_gen_ms"] * 1e3 / fs, grid, info["smem_bytes"]))
    if info["kernel_variant"] != 2:
        continue
    pall = prof.cpu().numpy().astype(np.float64)[8:-1, :16]
    p = pall[:, :, :14]      # core CTAs
    base = p[:, :, 0].min(axis=1, keepdims=True)[:, :, None]
    rel = p - base
    names = ["iter start", "L0 input ready", "L1 input ready", "L2 input ready", "L3 input ready", "L4 input ready", "L5 input ready",
             "layers done", "x1 published", "x2 published", "sampled", "next h0 published", "next partial done"]
    print("   helper-wait cycles in layer loop (thread 0 of each core CTA): mean %.0f" % p[:, :, 13].mean())
    per = np.diff(p[:, 0, 0]).mean()
    print("   step period %.0f ns" % per)
    for e, nme in enumerate(names):
        print("   %-20s mean %.0f ns  (min CTA %.0f, max CTA %.0f)" % (nme, rel[:, :, e].mean(), rel[:, :, e].mean(0).min(), rel[:, :, e].mean(0).max()))
    if pall.shape[2] >= 32 and pall[:, :, 16:32].max() > 0:
        print("   per-role stamps, ns since iter start of the earliest CTA (mean over core CTAs): warp 0 / 4 / 8 / 15")
        for ev, nme in ((13, "warp 0: x1 landed"), (9, "warp 0: x2 row published"), (15, "warp 0: x2 landed"), (16, "warp 0: score published"),
                        (17, "warp 0: scores landed"), (10, "warp 0: arg-max done")):
            print("      %-52s %6.0f" % (nme, (pall[:, :, ev:ev + 1] - base).mean()))
        for ev, nme in ((20, "leaves its layer section"), (24, "arg-max done")):
            v = pall[:, :, ev:ev + 4] - base
            print("      %-52s %s" % (nme, "  ".join("%6.0f" % x for x in v.mean(axis=(0, 1)))))
    # clock64 stamps of thread 0 (chain warp 0): layer 3 stage and the head, cycles
    c = pall
    lay = ["before wait", "woken", "h frags + taps in regs", "taps reduced", "gated", "published", "next taps issued"]
    print("   layer-3 stage of chain warp 0, cycles since 'before wait' (mean over core CTAs):")
    for i, nme in enumerate(lay):
        print("      %-26s %7.0f" % (nme, (c[:, :, 14 + i] - c[:, :, 14]).mean()))
    hd = [(31, "last layer output landed"), (21, "before x1 wait"), (22, "x1 landed"), (23, "out_1 row published"), (24, "x2 landed"),
          (25, "out_2 row published"), (26, "scores landed"), (27, "argmax done"), (28, "CTA barrier"), (29, "barriers re-armed"),
          (30, "next h0 published")]
    print("   head of warp 0, cycles since the last layer output landed:")
    for e, nme in hd:
        print("      %-26s %7.0f" % (nme, (c[:, :, e] - c[:, :, 31]).mean()))
