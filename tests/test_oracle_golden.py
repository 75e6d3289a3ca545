"""Pin the CPU oracle against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from _util import golden_names, load_golden, assert_close, large_names, load_large, head_errors
from oracle import swn_oracle as O


@pytest.mark.parametrize("name", golden_names())
def test_teacher_forced_heads_match_reference(name):
    g = load_golden(name)
    heads = O.teacher_forced(g["cfg"], g["sd"], g["aux"], g["samples"])
    assert_close(heads, g["heads"], name + " heads", rtol=2e-5)


@pytest.mark.parametrize("name", golden_names())
def test_free_running_generation_matches_reference(name):
    g = load_golden(name)
    cfg = g["cfg"]
    n = g["samples"].shape[1]
    smp, heads = O.generate(cfg, g["sd"], g["aux"], n, g["noise"], return_heads=True)
    assert_close(heads, g["heads"], name + " heads", rtol=2e-5)
    if cfg.kind == "dswnv":
        assert smp.dtype == np.int64
        assert np.array_equal(smp, g["samples"])
    else:
        assert smp.dtype == np.float32
        np.testing.assert_allclose(smp, g["samples"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", golden_names())
def test_forced_generation_equals_teacher_forced(name):
    g = load_golden(name)
    cfg = g["cfg"]
    n = g["samples"].shape[1]
    smp, heads = O.generate(cfg, g["sd"], g["aux"], n, g["noise"], forced=g["samples"], return_heads=True)
    assert_close(heads, O.teacher_forced(cfg, g["sd"], g["aux"], g["samples"]), name, rtol=2e-5)
    if cfg.kind == "dswnv":
        assert np.array_equal(smp, g["samples"])


def test_gumbel_form_equals_multinomial_form():
    g = load_golden("dswnv_base")
    logits = g["heads"].reshape(-1, 256)
    q = g["noise"].reshape(-1, 256)
    a = O.softmax_sample(torch.from_numpy(logits), torch.from_numpy(q)).numpy()
    b = O.gumbel_argmax(logits, q)
    assert np.array_equal(a, b)
    assert np.array_equal(a, g["samples"].reshape(-1))


def test_mu_law_roundtrip_and_seed():
    assert int(O.encode_mu_law(np.zeros(1))[0]) == 128          # decode_dswnv_softmax.py seed
    y = np.arange(256)
    x = O.decode_mu_law(y)
    assert np.abs(x).max() < 1.03                     # decode uses (y-0.5): dswnv.py:45
    assert np.all(np.diff(x) > 0)
    back = O.encode_mu_law(x)
    assert np.abs(back - y).max() <= 1


def test_truncation_contract():
    g = load_golden("cswnv_seg5_lpc4_wav")
    outs = O.split_outputs(g["samples"], g["n_samples_list"])
    assert [len(o) for o in outs] == g["n_samples_list"]


@pytest.mark.parametrize("name", large_names())
def test_oracle_matches_recipe_size_reference_fixture(name):
    """tests/golden_large (oracle/build_ref.py): the UNMODIFIED reference at the size the benchmark runs (H=192 seg 5 /
    H=256, k=7, L=6, R=690).  Pins the oracle - and the weights rebuilt from the seed - at that size."""
    g = load_large(name)
    heads = O.teacher_forced(g["cfg"], g["sd"], g["aux"], g["samples"])
    errs = head_errors(heads, g["heads"], g["kind"], g["kw"])
    assert max(errs.values()) <= 2e-5, errs


@pytest.mark.parametrize("name", ["large_cswnv_seg5_b2", "large_dswnv_b1"])
def test_oracle_free_running_matches_recipe_size_reference_fixture(name):
    g = load_large(name)
    n = g["samples"].shape[1]
    smp = O.generate(g["cfg"], g["sd"], g["aux"], n, g["noise"])
    if g["kind"] == "dswnv":
        assert np.array_equal(smp, g["samples"])
    else:
        np.testing.assert_allclose(smp, g["samples"], rtol=0, atol=1e-5)


def test_reference_runner_reproduces_a_committed_fixture():
    """oracle/ref_runner.py (the CPU arm of bench.py) drives the unmodified reference modules - from oracle/_ref/ on the
    GPU box, from /root/reference in the build container - and reproduces a fixture that make_golden.py wrote."""
    from oracle import ref_runner as RR
    if not RR.available():
        pytest.skip("no reference modules here (oracle/_ref is written by __graft_entry__.build() in the build container)")
    g = load_golden("dswnv_base")
    model = RR.build_model("dswnv", g["kw"], g["sd"])
    B, N = g["samples"].shape
    out, heads, dt = RR.generate("dswnv", g["kw"], model, g["aux"], [N] * B, noise=g["noise"], capture_heads=True)
    assert np.array_equal(np.stack(out), g["samples"])
    assert float(np.abs(heads - g["heads"]).max()) <= 1e-6
    gc = load_golden("cswnv_seg5_lpc4_wav")
    mc = RR.build_model("cswnv", gc["kw"], gc["sd"])
    B, N = gc["samples"].shape
    out, heads, dt = RR.generate("cswnv", gc["kw"], mc, gc["aux"], [N] * B, noise=gc["noise"], capture_heads=True)
    assert float(np.abs(np.stack(out) - gc["samples"]).max()) <= 1e-6
