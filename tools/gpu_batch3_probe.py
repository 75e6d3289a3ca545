"""Throughput probe of the pipelined batch kernel (variant 3) on the GPU box: sweeps the slot size."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shallow_wavenet_b200 import synth

frames = int(os.environ.get("FRAMES", "20"))
cases = []
for B, tiles in ((64, (4, 8, 16)), (256, (8, 16))):
    for t in tiles:
        cases.append(("cswnv", 22050, dict(seg=5, lpc=4), B, t))
for B, tiles in ((128, (4, 8)),):
    for t in tiles:
        cases.append(("dswnv", 24000, dict(), B, t))
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] == sys.argv[1]]
for kind, fs, over, B, tile in cases:
    kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
    kw.update(over)
    model = synth.make_model(kind, **kw).cuda()
    aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
    eng = model.engine()
    n = frames * kw["upsampling_factor"]
    try:
        for rep in range(2):
            eng.condition(aux)
            out, _ = eng.generate(n, seed=rep, batch_tile=tile, kernel_sel=int(os.environ.get("ENGINE", "3")))
        info = eng.info
        seg = info["eff_seg"]
        steps = n // seg + info["receptive_field"] // seg + 1
        print("%s B=%3d tile %2d: variant %d slots %3d grid %d heads %d smem %6d | kernel %8.2f ms cond %6.2f ms | %7.2f us/iteration | %9.0f samples/s total, %7.1fx RT aggregate" % (
            kind, B, info["batch_tile"], info["kernel_variant"], info["n_slots"], info["grid"], info["head_ctas"], info["smem_bytes"],
            info["last_gen_ms"], info["last_cond_ms"], info["last_gen_ms"] * 1e3 / steps,
            B * (n + info["receptive_field"]) / info["last_gen_ms"] * 1e3, B * (n + info["receptive_field"]) / info["last_gen_ms"] * 1e3 / fs), flush=True)
    except Exception as e:
        print("%s B=%d tile %d FAILED: %s" % (kind, B, tile, e), flush=True)
    del model, eng
    torch.cuda.empty_cache()
