"""Ad-hoc GPU parity report (prints errors instead of asserting). Run on the GPU box."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import shallow_wavenet_b200 as S
from _util import golden_names, load_golden

only = sys.argv[1:]
for name in golden_names():
    if only and name not in only:
        continue
    g = load_golden(name)
    cls = S.CSWNV if g["kind"] == "cswnv" else S.DSWNV
    m = cls(**g["kw"])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in g["sd"].items()})
    m.cuda().eval()
    B, N = g["samples"].shape
    seed_audio = torch.zeros(B, g["cfg"].eff_seg) if g["kind"] == "cswnv" else torch.full((B, 1), 128, dtype=torch.long)
    t = time.time()
    try:
        outs, heads = m.batch_fast_generate(seed_audio, torch.from_numpy(g["aux"]), [N] * B, noise=g["noise"],
                                            forced=g["samples"], return_heads=True)
    except Exception as e:
        print("%-26s FAILED: %s" % (name, e)); continue
    smp = np.stack(outs)
    herr = np.abs(heads - g["heads"])
    scale = np.abs(g["heads"]).max(-1, keepdims=True)
    rel = (herr / (scale + 0.1)).max()
    if g["kind"] == "dswnv":
        serr = int((smp != g["samples"]).sum())
    else:
        serr = float(np.abs(smp - g["samples"]).max())
    print("%-26s forced: head max abs %.3e rel %.3e  first-step err %.3e  sample err %s  (%.2fs) info=%s" % (
        name, herr.max(), rel, herr[:, 0].max(), serr, time.time() - t,
        {k: m.engine().info[k] for k in ("kernel_variant", "grid", "batch_tile", "n_slots", "smem_bytes", "last_gen_ms")}))
    try:
        outs2 = m.batch_fast_generate(seed_audio, torch.from_numpy(g["aux"]), [N] * B, noise=g["noise"])
        smp2 = np.stack(outs2)
        if g["kind"] == "dswnv":
            print("    free-running mismatches: %d / %d" % (int((smp2 != g["samples"]).sum()), smp2.size))
        else:
            print("    free-running max abs err: %.3e" % float(np.abs(smp2 - g["samples"]).max()))
    except Exception as e:
        print("    free-running FAILED: %s" % e)
