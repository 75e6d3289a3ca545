"""Training-time / evaluation ``forward``: ``CSWNV.forward(aux, audio)`` /
``DSWNV.forward(audio, aux)`` of the reference (cswnv_shift1.py:191-267, dswnv.py:250-276), computed time-parallel on
the GPU by tcgen05 (tf32 x 3) GEMMs (csrc/swn_fwd.cu).  tests/golden_forward/*.npz are outputs of the UNMODIFIED
reference modules (tests/golden/make_forward_golden.py).

Tolerance: every returned tensor, scaled per row, |a-b| <= FWD_TOL * (max|row of the reference| + 0.05) with
FWD_TOL = 1e-4, the north-star tolerance (three tf32 passes through 6 layers + 3 head GEMMs: measured <= 3.5e-5).  CPU: the oracle's
restatement of forward against the same fixtures at 2e-5."""
import json
import os

import numpy as np
import pytest
import torch

from _util import forward_names, load_forward, scaled_err
from oracle import swn_oracle as O

FWD_TOL = 1e-4        # the north-star tolerance; measured on a B200: <= 3.5e-5 (profiles/r02_parity_errors_forward.txt)


def _ref_keys(g):
    return sorted(g["ref"])


@pytest.mark.parametrize("name", forward_names())
def test_oracle_forward_matches_reference_fixture(name):
    g = load_forward(name)
    r = O.forward_full(g["cfg"], g["sd"], g["aux"], g["audio"], clip=True)
    for k, ref in g["ref"].items():
        mine = r["out"] if k == "logits" else r[k]
        if g["kind"] == "cswnv" and g["cfg"].seg == 1:
            mine = mine.reshape(mine.shape[0], -1)
        assert scaled_err(mine, ref) <= 2e-5, (name, k)


def _model(g):
    import shallow_wavenet_b200 as S
    m = (S.CSWNV if g["kind"] == "cswnv" else S.DSWNV)(**g["kw"])
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in g["sd"].items()}, strict=True)
    return m.cuda().eval()


def _log(what, errs):
    log = os.environ.get("SWN_PARITY_LOG")
    if log:
        with open(log, "a") as f:
            f.write(json.dumps({"what": what, "tol": FWD_TOL, "errors": errs}) + "\n")


@pytest.mark.gpu
@pytest.mark.parametrize("name", forward_names())
def test_forward_matches_reference_fixture(name):
    g = load_forward(name)
    m = _model(g)
    aux = torch.from_numpy(g["aux"])
    errs = {}
    if g["kind"] == "cswnv":
        audio = torch.from_numpy(g["audio"]).unsqueeze(1)                 # [B, 1, T_a] as the training script passes it
        out = m(aux, audio)
        names = ("mu", "b", "log_b", "a") if g["cfg"].lpc > 0 else ("mu", "b", "log_b")
        assert len(out) == len(names)
        for n, t in zip(names, out):
            assert t.is_cuda and tuple(t.shape) == g["ref"][n].shape, (n, tuple(t.shape), g["ref"][n].shape)
            errs[n] = scaled_err(t.cpu().numpy(), g["ref"][n])
        outc = m(aux, audio, clip=True)
        namesc = ("mu", "b", "b_clip", "log_b_clip", "a") if g["cfg"].lpc > 0 else ("mu", "b", "b_clip", "log_b_clip")
        assert len(outc) == len(namesc)
        for n, t in zip(namesc, outc):
            errs["clip." + n] = scaled_err(t.cpu().numpy(), g["ref"][n])
    else:
        Q = g["kw"]["n_quantize"]
        onehot = torch.nn.functional.one_hot(torch.from_numpy(g["audio"]), Q).float().transpose(1, 2)   # the reference's input
        logits = m(onehot, aux)
        assert logits.is_cuda and tuple(logits.shape) == g["ref"]["logits"].shape
        errs["logits"] = scaled_err(logits.cpu().numpy(), g["ref"]["logits"])
        again = m(torch.from_numpy(g["audio"]), aux)                       # class indices instead of the one-hot tensor
        assert torch.equal(again, logits)
    _log("forward " + name, errs)
    assert max(errs.values()) <= FWD_TOL, errs
    assert m.engine().info["kernel_variant"] == 5


@pytest.mark.gpu
@pytest.mark.parametrize("kind,fs,B,frames,over", [("cswnv", 22050, 3, 10, dict(seg=5, lpc=4)), ("cswnv", 22050, 2, 6, dict(seg=1, lpc=4)),
                                                   ("cswnv", 22050, 1, 5, dict(seg=10, lpc=0)), ("dswnv", 24000, 2, 8, dict())])
def test_forward_recipe_size_several_tiles_vs_oracle(kind, fs, B, frames, over):
    """Recipe size, utterances longer than one 256-step tile and than the longest tap delay (6 * 49 = 294 steps)."""
    from shallow_wavenet_b200 import synth
    kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
    kw.update(over)
    model = synth.make_model(kind, seed=5, **kw)
    sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    cfg = O.SWNConfig(kind=kind, **kw)
    aux = synth.synth_aux(kw["n_aux"], frames, B, seed=77)
    rng = np.random.default_rng(3)
    seg = cfg.eff_seg
    Ta = frames * kw["upsampling_factor"] - seg
    model = model.cuda()
    if kind == "cswnv":
        audio = np.clip(0.3 * rng.standard_normal((B, Ta)), -1, 1).astype(np.float32)
        out = model(torch.from_numpy(aux), torch.from_numpy(audio))
        ref = O.forward_full(cfg, sd, aux, audio)
        names = ("mu", "b", "log_b", "a") if cfg.lpc > 0 else ("mu", "b", "log_b")
        errs = {n: scaled_err(t.cpu().numpy(), ref[n]) for n, t in zip(names, out)}
    else:
        audio = rng.integers(0, 256, (B, Ta)).astype(np.int64)
        errs = {"logits": scaled_err(model(torch.from_numpy(audio), torch.from_numpy(aux)).cpu().numpy(),
                                     O.forward_full(cfg, sd, aux, audio)["out"])}
    _log("forward recipe %s B=%d T=%d" % (kind, B, Ta - seg + 1), errs)
    assert max(errs.values()) <= FWD_TOL, errs


@pytest.mark.gpu
def test_forward_argument_checks():
    from shallow_wavenet_b200._abi import SwnError
    g = load_forward("fwd_dswnv_base")
    m = _model(g)
    aux = torch.from_numpy(g["aux"])
    with pytest.raises(SwnError, match="audio length"):
        m(torch.from_numpy(g["audio"][:, :-1]), aux)
    soft = torch.full((g["audio"].shape[0], 256, g["audio"].shape[1]), 1.0 / 256)
    with pytest.raises(SwnError, match="not one-hot"):
        m(soft, aux)
