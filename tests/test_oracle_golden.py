"""Pin the CPU oracle against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from _util import golden_names, load_golden, assert_close
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
