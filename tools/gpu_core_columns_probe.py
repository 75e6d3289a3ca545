"""Step time of the cluster kernel by the number of utterance columns it carries (NC = 1, 2, 4 instantiations), DSWNV recipe
size and CSWNV seg=1, next to what `kernel_sel=0` (the automatic dispatch) picks for B = 1..8."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from shallow_wavenet_b200 import synth  # noqa: E402

frames = 60
for kind, fs, over in (("dswnv", 16000, dict()), ("cswnv", 22050, dict(seg=1, lpc=4))):
    kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
    kw.update(over)
    model = synth.make_model(kind, **kw).cuda()
    eng = model.engine()
    n = frames * kw["upsampling_factor"]
    for B, tile in ((1, 1), (2, 2), (3, 4), (4, 4)):
        aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
        try:
            for rep in range(3):
                eng.condition(aux)
                eng.generate(n, seed=rep, kernel_sel=2, batch_tile=tile)
            i = eng.info
            steps = n + i["receptive_field"] + 1
            print("%s cluster kernel B=%d columns=%d: %8.3f us/step %9.0f samples/s" % (
                kind, B, tile, i["last_gen_ms"] * 1e3 / steps, B * n / i["last_gen_ms"] * 1e3), flush=True)
        except Exception as e:
            print("%s cluster kernel B=%d columns=%d: n/a (%s)" % (kind, B, tile, str(e)[:80]), flush=True)
    for B in (1, 2, 3, 4, 5, 6, 7, 8):
        aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
        for rep in range(3):
            eng.condition(aux)
            eng.generate(n, seed=rep)
        i = eng.info
        print("%s auto B=%d: variant %d tile %d launches/slots %d  %9.0f samples/s" % (
            kind, B, i["kernel_variant"], i["batch_tile"], i["n_slots"], B * n / i["last_gen_ms"] * 1e3), flush=True)
