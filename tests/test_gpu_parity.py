"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the Python
mirror of the reference API (a ctypes caller of the C ABI), against
  * the golden fixtures produced by the unmodified reference (tests/golden), and
  * the CPU oracle on seeded inputs at recipe sizes,
plus size-independent properties at BASELINE.json's full sizes.

Tolerances (north-star: "within a stated fp32 tolerance, e.g. 1e-4 relative"):
  head values, PER COMPONENT (logits | Laplace loc | pre-sigmoid scale | LP coefficients):
      |a-b| <= tol * (max|component of the row| + 0.05), tol = 2e-5 on the fp32 FFMA kernels and
      3e-5 on the tensor-core cells (three tf32 passes)
  Laplace samples given the same head, history and noise: |a-b| <= 2e-5 absolute (range [-1,1])
  mu-law indices: bit-exact given the logits (checked against the argmax recomputed on the host
  from the kernel's own logits and noise, and against the reference's samples on the fixtures).
"""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from _util import golden_names, load_golden, assert_close, head_errors, large_names, load_large
from oracle import swn_oracle as O

pytestmark = pytest.mark.gpu

HEAD_TOL = 2e-5            # fp32 FFMA kernels (variants 1, 2, 4): measured <= 5.2e-6
HEAD_TOL_TF32X3 = 3e-5     # tensor-core cells (variant 3: W_hi*X_hi + W_lo*X_hi + W_hi*X_lo in tf32): measured <= 8e-6
                           # (profiles/r02_parity_errors.txt); the north-star allows 1e-4
SAMPLE_ATOL = 2e-5


def _S():
    import shallow_wavenet_b200 as S
    return S


def _model_from_golden(g):
    S = _S()
    cls = S.CSWNV if g["kind"] == "cswnv" else S.DSWNV
    m = cls(**g["kw"])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in g["sd"].items()}, strict=True)
    return m.cuda().eval()


def _seed_audio(kind, B, seg):
    return torch.zeros(B, seg) if kind == "cswnv" else torch.full((B, 1), 128, dtype=torch.long)


def _head_close(a, b, what, g=None, kind=None, kw=None, tol=None):
    """Per-component parity of out_2 rows (loc / pre-sigmoid scale / LP coefficients separately, or the 256 logits):
    max|a-b| <= tol * (max|component of the reference row| + 0.05).  tol defaults to HEAD_TOL (fp32 FFMA kernels);
    the tensor-core cells (three tf32 passes) are held to HEAD_TOL_TF32X3.  Every call appends its measured errors to
    $SWN_PARITY_LOG (if set) so that the tolerances stay calibrated against what the kernels actually achieve."""
    if g is not None:
        kind, kw = g["kind"], g["kw"]
    tol = HEAD_TOL if tol is None else tol
    errs = head_errors(a, b, kind, kw)
    log = os.environ.get("SWN_PARITY_LOG")
    if log:
        with open(log, "a") as f:
            f.write(json.dumps({"what": what, "tol": tol, "errors": errs}) + "\n")
    for comp, err in errs.items():
        assert err <= tol, "%s: %s scaled head error %.3e > %.1e" % (what, comp, err, tol)
    return max(errs.values())


def _neg_log_q(u):
    """-log q of the in-kernel sampler for Philox uniforms u in [0, 1): q = max(-log(1 - u), 2^-24) (swn_gen.cuh softmax_q;
    the floor keeps u == 0 from producing an infinite Gumbel score, as torch's exponential_ does)."""
    q = np.maximum(-np.log1p(-u.astype(np.float64)), 5.9604645e-8)
    return (-np.log(q)).astype(np.float32)


def _tol(model):
    """Tolerance by the kernel that produced the heads (info of the model's last launch)."""
    return HEAD_TOL_TF32X3 if model.engine().info["kernel_variant"] == 3 else HEAD_TOL


def test_native_library_is_loaded():
    from shallow_wavenet_b200 import _abi
    _abi.load()
    maps = open("/proc/self/maps").read()
    assert "libswn_b200.so" in maps


@pytest.mark.parametrize("name", golden_names())
def test_teacher_forced_matches_reference_fixture(name):
    g = load_golden(name)
    m = _model_from_golden(g)
    B, N = g["samples"].shape
    outs, heads = m.batch_fast_generate(_seed_audio(g["kind"], B, g["cfg"].eff_seg), torch.from_numpy(g["aux"]),
                                        [N] * B, noise=g["noise"], forced=g["samples"], return_heads=True)
    _head_close(heads, g["heads"], name, g, tol=_tol(m))
    smp = np.stack(outs)
    if g["kind"] == "dswnv":
        assert smp.dtype == np.int64 and np.array_equal(smp, g["samples"])
    else:
        assert smp.dtype == np.float32
        assert float(np.abs(smp - g["samples"]).max()) <= SAMPLE_ATOL


@pytest.mark.parametrize("name", golden_names())
def test_swn_teacher_forced_entry_point_matches_reference_fixture(name):
    """swn_teacher_forced: the reference's own samples in, the reference's per-step out_2 rows out."""
    g = load_golden(name)
    m = _model_from_golden(g)
    eng = m.engine()
    eng.condition(torch.from_numpy(g["aux"]).cuda())
    heads = eng.teacher_forced(g["samples"]).cpu().numpy()
    _head_close(heads, g["heads"], name, g, tol=_tol(m))


@pytest.mark.parametrize("name", golden_names())
def test_free_running_matches_reference_fixture(name):
    g = load_golden(name)
    m = _model_from_golden(g)
    B, N = g["samples"].shape
    outs = m.batch_fast_generate(_seed_audio(g["kind"], B, g["cfg"].eff_seg), torch.from_numpy(g["aux"]),
                                 g["n_samples_list"], noise=g["noise"])
    assert [len(o) for o in outs] == g["n_samples_list"]         # per-utterance truncation (cswnv:426-428)
    for b, o in enumerate(outs):
        ref = g["samples"][b, :len(o)]
        if g["kind"] == "dswnv":
            assert np.array_equal(o, ref)
        else:
            assert float(np.abs(o - ref).max()) <= 5e-5          # 60 fed-back steps of fp32 reordering


# ----------------------------------------------------------------------------- recipe-size reference fixtures
@pytest.mark.parametrize("name", large_names())
def test_recipe_size_matches_reference_fixture(name):
    """tests/golden_large: outputs of the UNMODIFIED reference at the size the benchmark runs (H=192 seg 5 / H=256,
    k=7, L=6, R=690), written by oracle/build_ref.py.  Teacher-forced heads and samples, then free running on the
    reference's own noise, through whichever kernel the plan selects for that batch size (cluster kernel at B<=2 and for seg 5 at B<=4,
    the batch kernel from 8 utterances)."""
    g = load_large(name)
    m = g["model"].cuda().eval()
    B, N = g["samples"].shape
    seg = g["cfg"].eff_seg
    seed_a = _seed_audio(g["kind"], B, seg)
    outs, heads = m.batch_fast_generate(seed_a, torch.from_numpy(g["aux"]), [N] * B, noise=g["noise"],
                                        forced=g["samples"], return_heads=True)
    variant = m.engine().info["kernel_variant"]
    _head_close(heads, g["heads"], "%s teacher-forced v%d" % (name, variant), g, tol=_tol(m))
    smp = np.stack(outs)
    if g["kind"] == "dswnv":
        assert np.array_equal(smp, g["samples"])
    else:
        assert float(np.abs(smp - g["samples"]).max()) <= SAMPLE_ATOL
    outs = m.batch_fast_generate(seed_a, torch.from_numpy(g["aux"]), g["n_samples_list"], noise=g["noise"])
    assert [len(o) for o in outs] == [(n // seg) * seg for n in g["n_samples_list"]]
    for b, o in enumerate(outs):
        ref = g["samples"][b, :len(o)]
        if g["kind"] == "dswnv":
            assert np.array_equal(o, ref), (name, b, int((o != ref).sum()))
        else:
            assert float(np.abs(o - ref).max()) <= 5e-5


@pytest.mark.parametrize("name,sel", [("large_dswnv_b2", 1), ("large_dswnv_b8", 4), ("large_cswnv_seg5_b8", 4),
                                      ("large_cswnv_seg1_b1", 1), ("large_cswnv_seg5_b2", 3)])
def test_recipe_size_reference_fixture_on_the_other_kernels(name, sel):
    g = load_large(name)
    m = g["model"].cuda().eval()
    B, N = g["samples"].shape
    outs, heads = m.batch_fast_generate(_seed_audio(g["kind"], B, g["cfg"].eff_seg), torch.from_numpy(g["aux"]), [N] * B,
                                        noise=g["noise"], forced=g["samples"], return_heads=True, kernel_sel=sel)
    assert m.engine().info["kernel_variant"] == sel
    _head_close(heads, g["heads"], "%s teacher-forced v%d" % (name, sel), g, tol=_tol(m))
    if g["kind"] == "dswnv":
        assert np.array_equal(np.stack(outs), g["samples"])
    else:
        assert float(np.abs(np.stack(outs) - g["samples"]).max()) <= SAMPLE_ATOL


def _recipe_case(kind, fs, B, frames, seed, **over):
    from shallow_wavenet_b200 import synth
    kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
    kw.update(over)
    model = synth.make_model(kind, seed=seed, **kw)
    aux = synth.synth_aux(kw["n_aux"], frames, B, seed=seed)
    cfg = O.SWNConfig(kind=kind, **kw)
    sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    return model.cuda(), cfg, sd, aux, kw


@pytest.mark.parametrize("seg,lpc,B", [(5, 4, 1), (5, 4, 3), (1, 4, 2), (1, 4, 3), (1, 0, 4), (2, 0, 1), (10, 4, 1), (3, 4, 2), (4, 2, 1), (5, 4, 5)])
def test_recipe_size_cswnv_teacher_forced_vs_oracle(seg, lpc, B):
    model, cfg, sd, aux, kw = _recipe_case("cswnv", 22050, B, 2, 7, seg=seg, lpc=lpc)
    N = 2 * kw["upsampling_factor"]
    rng = np.random.default_rng(5)
    forced = np.clip(0.3 * rng.standard_normal((B, N)), -1, 1).astype(np.float32)
    noise = rng.uniform(-0.4999, 0.5, (B, N)).astype(np.float32)
    n_gen = (N // seg) * seg                        # (seg 3: 219 of the 220 samples, as the reference's loop count)
    forced, noise = forced[:, :n_gen], noise[:, :n_gen]
    outs, heads = model.batch_fast_generate(torch.zeros(B, seg), torch.from_numpy(aux), [N] * B, noise=noise,
                                            forced=forced, return_heads=True)
    ref_heads = O.teacher_forced(cfg, sd, aux, forced)
    _head_close(heads, ref_heads, "cswnv seg%d B=%d" % (seg, B), kind="cswnv", kw=kw, tol=_tol(model))
    # sampler given the oracle's view of the same heads / history / noise
    lp = torch.zeros(B, lpc) if lpc else None
    ref = np.zeros((B, n_gen), np.float32)
    ft = torch.from_numpy(forced)
    for i in range(n_gen // seg):
        s, _ = O.laplace_sample(cfg, torch.from_numpy(ref_heads[:, i]), lp, torch.from_numpy(noise[:, i * seg:(i + 1) * seg]))
        ref[:, i * seg:(i + 1) * seg] = s.numpy()
        if lpc:
            lp = torch.cat((lp, ft[:, i * seg:(i + 1) * seg]), 1)[:, -lpc:]
    assert float(np.abs(np.stack(outs) - ref).max()) <= SAMPLE_ATOL
    if 2 <= seg <= 5 and B <= 4:
        # the seg cluster kernel, one utterance per core cluster: the Philox draws of utterance b use its global index b
        info = model.engine().info
        assert info["kernel_variant"] == 2 and 1 <= info["batch_tile"] <= B and info["n_slots"] == B, info
        from shallow_wavenet_b200.engine import philox_uniform
        eps = np.stack([np.float32(-0.4999) + np.float32(0.9999) * philox_uniform(9, "cswnv", b, 0, n_gen) for b in range(B)])
        a = np.stack(model.batch_fast_generate(torch.zeros(B, seg), torch.from_numpy(aux), [N] * B, seed=9))
        b_ = np.stack(model.batch_fast_generate(torch.zeros(B, seg), torch.from_numpy(aux), [N] * B, noise=eps.astype(np.float32)))
        assert float(np.abs(a - b_).max()) <= 1e-6
    if seg == 1 and B >= 3:
        # launches of two through the cluster kernel; the Philox draws of a later launch use the global utterance index
        info = model.engine().info
        assert info["kernel_variant"] == 2 and info["batch_tile"] == 2 and info["n_slots"] == -(-B // 2), info
        from shallow_wavenet_b200.engine import philox_uniform
        eps = np.stack([np.float32(-0.4999) + np.float32(0.9999) * philox_uniform(9, "cswnv", b, 0, N) for b in range(B)])
        a = np.stack(model.batch_fast_generate(torch.zeros(B, seg), torch.from_numpy(aux), [N] * B, seed=9))
        b_ = np.stack(model.batch_fast_generate(torch.zeros(B, seg), torch.from_numpy(aux), [N] * B, noise=eps.astype(np.float32)))
        assert float(np.abs(a - b_).max()) <= 1e-6


@pytest.mark.parametrize("fs,B", [(16000, 1), (24000, 3)])
def test_recipe_size_dswnv_teacher_forced_vs_oracle(fs, B):
    model, cfg, sd, aux, kw = _recipe_case("dswnv", fs, B, 2, 9)
    N = 2 * kw["upsampling_factor"]
    rng = np.random.default_rng(6)
    forced = rng.integers(0, 256, (B, N)).astype(np.int64)
    q = rng.exponential(1.0, (B, N, 256)).astype(np.float32)
    outs, heads = model.batch_fast_generate(torch.full((B, 1), 128, dtype=torch.long), torch.from_numpy(aux), [N] * B,
                                            noise=q, forced=forced, return_heads=True)
    ref_heads = O.teacher_forced(cfg, sd, aux, forced)
    _head_close(heads, ref_heads, "dswnv %d B=%d" % (fs, B), kind="dswnv", kw=kw, tol=_tol(model))
    smp = np.stack(outs)
    mine = O.gumbel_argmax(heads.reshape(-1, 256), q.reshape(-1, 256)).reshape(B, N)
    assert np.array_equal(smp, mine)                              # bit-exact given the kernel's logits
    ref_idx = O.softmax_sample(torch.from_numpy(ref_heads.reshape(-1, 256)), torch.from_numpy(q.reshape(-1, 256))).numpy()
    assert (smp.reshape(-1) != ref_idx).mean() < 2e-3             # reference form on reference logits: near-ties only


@pytest.mark.parametrize("B,tile", [(5, 2), (5, 8), (11, 4), (3, 1)])
def test_batch_tiles_and_pipeline_slots_dswnv(B, tile):
    g = load_golden("dswnv_base")
    m = _model_from_golden(g)
    rng = np.random.default_rng(B)
    Tf = 4
    U = g["kw"]["upsampling_factor"]
    N = Tf * U
    aux = rng.standard_normal((B, g["kw"]["n_aux"], Tf)).astype(np.float32)
    q = rng.exponential(1.0, (B, N, 256)).astype(np.float32)
    outs, heads = m.batch_fast_generate(torch.full((B, 1), 128, dtype=torch.long), torch.from_numpy(aux), [N] * B,
                                        noise=q, return_heads=True, batch_tile=tile)
    info = m.engine().info
    assert info["batch_tile"] == tile and info["n_slots"] == -(-B // tile)
    ref_smp, ref_heads = O.generate(g["cfg"], g["sd"], aux, N, q, return_heads=True)
    assert np.array_equal(np.stack(outs), ref_smp)
    _head_close(heads, ref_heads, "B=%d tile=%d" % (B, tile), g, tol=_tol(m))


@pytest.mark.parametrize("B,tile", [(3, 1), (3, 2), (5, 2)])
def test_batch_tiles_and_pipeline_slots_cswnv(B, tile):
    g = load_golden("cswnv_seg5_lpc4_wav")
    m = _model_from_golden(g)
    rng = np.random.default_rng(B)
    Tf, U = 5, g["kw"]["upsampling_factor"]
    N = Tf * U
    aux = rng.standard_normal((B, g["kw"]["n_aux"], Tf)).astype(np.float32)
    eps = rng.uniform(-0.4999, 0.5, (B, N)).astype(np.float32)
    outs, heads = m.batch_fast_generate(torch.zeros(B, 5), torch.from_numpy(aux), [N] * B, noise=eps,
                                        return_heads=True, batch_tile=tile)
    ref_smp, ref_heads = O.generate(g["cfg"], g["sd"], aux, N, eps, return_heads=True)
    _head_close(heads, ref_heads, "B=%d tile=%d" % (B, tile), g, tol=_tol(m))
    assert float(np.abs(np.stack(outs) - ref_smp).max()) <= 5e-5


def test_non_default_seed_waveform():
    g = load_golden("cswnv_seg5_lpc4_wav")
    m = _model_from_golden(g)
    rng = np.random.default_rng(1)
    B, N = 2, 30
    aux = g["aux"][:B]
    seed_wav = (0.2 * rng.standard_normal((B, 5))).astype(np.float32)
    eps = rng.uniform(-0.4999, 0.5, (B, N)).astype(np.float32)
    outs, heads = m.batch_fast_generate(torch.from_numpy(seed_wav), torch.from_numpy(aux), [N] * B, noise=eps,
                                        return_heads=True)
    # oracle with the same seed: patch the padded audio through the module-level helper
    orig = O._audio_full

    def seeded(cfg, gen, Bn, n):
        a = orig(cfg, gen, Bn, n)
        a[:, cfg.receptive_field: cfg.receptive_field + cfg.seg] = torch.from_numpy(seed_wav)
        return a
    O._audio_full = seeded
    try:
        ref_heads = O.teacher_forced(g["cfg"], g["sd"], aux, np.stack(outs))
    finally:
        O._audio_full = orig
    _head_close(heads, ref_heads, "seeded", g, tol=_tol(m))
    assert np.abs(heads - O.teacher_forced(g["cfg"], g["sd"], aux, np.stack(outs))).max() > 1e-4   # the seed matters


def test_side_by_side_core_clusters_with_seed_waveform():
    """Five utterances of a one-sample-per-step CSWNV through the automatic choice: launches of 3 + 2, two core clusters
    side by side (two utterances + one, one + one), each cluster with its own seed sample, noise, conditioning table and
    outputs (CoreParams::cl[]): per-step heads against the oracle run with the same seed waveform."""
    g = load_golden("cswnv_seg1_lpc0")
    m = _model_from_golden(g)
    kw = g["kw"]
    rng = np.random.default_rng(5)
    B, T = 5, g["aux"].shape[2]
    N = T * kw["upsampling_factor"]
    aux = rng.standard_normal((B, kw["n_aux"], T)).astype(np.float32)
    seed_wav = (0.2 * rng.standard_normal((B, 1))).astype(np.float32)
    eps = rng.uniform(-0.4999, 0.5, (B, N)).astype(np.float32)
    outs, heads = m.batch_fast_generate(torch.from_numpy(seed_wav), torch.from_numpy(aux), [N] * B, noise=eps,
                                        return_heads=True)
    info = m.engine().info
    assert info["kernel_variant"] == 2 and info["batch_tile"] == 2 and info["n_slots"] == 3, info
    orig = O._audio_full

    def seeded(cfg, gen, Bn, n):
        a = orig(cfg, gen, Bn, n)
        a[:, cfg.receptive_field: cfg.receptive_field + cfg.seg] = torch.from_numpy(seed_wav)
        return a
    O._audio_full = seeded
    try:
        ref_heads = O.teacher_forced(g["cfg"], g["sd"], aux, np.stack(outs))
    finally:
        O._audio_full = orig
    _head_close(heads, ref_heads, "side-by-side clusters, seeded", g, tol=_tol(m))


def test_philox_stream_is_reproducible_and_matches_host_replica():
    from shallow_wavenet_b200 import engine as E
    g = load_golden("dswnv_base")
    m = _model_from_golden(g)
    B = 2
    N = 40
    aux = torch.from_numpy(g["aux"])
    seed_a = torch.full((B, 1), 128, dtype=torch.long)
    a = np.stack(m.batch_fast_generate(seed_a, aux, [N] * B, seed=77))
    b = np.stack(m.batch_fast_generate(seed_a, aux, [N] * B, seed=77))
    c = np.stack(m.batch_fast_generate(seed_a, aux, [N] * B, seed=78))
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    u = np.stack([E.philox_uniform(77, "dswnv", bb, 0, N) for bb in range(B)])          # [B,N,256] in [0,1)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    q = np.maximum(-np.log1p(-u.astype(np.float64)), 5.9604645e-8).astype(np.float32)
    d = np.stack(m.batch_fast_generate(seed_a, aux, [N] * B, noise=q))
    assert (a != d).mean() <= 0.05                                 # same stream up to logf rounding of q

    gc = load_golden("cswnv_seg5_lpc4_wav")
    mc = _model_from_golden(gc)
    Bc, Nc = gc["samples"].shape
    auxc = torch.from_numpy(gc["aux"])
    x = np.stack(mc.batch_fast_generate(torch.zeros(Bc, 5), auxc, [Nc] * Bc, seed=5))
    uc = np.stack([E.philox_uniform(5, "cswnv", bb, 0, Nc) for bb in range(Bc)])
    eps = (np.float32(-0.4999) + np.float32(0.9999) * uc).astype(np.float32)
    y = np.stack(mc.batch_fast_generate(torch.zeros(Bc, 5), auxc, [Nc] * Bc, noise=eps))
    assert float(np.abs(x - y).max()) <= 1e-6
    assert np.abs(x).max() <= 1.0


@pytest.mark.parametrize("B", [3, 4, 7])
def test_small_batches_go_through_the_cluster_kernel_two_at_a_time(B):
    """3-7 utterances (DSWNV): launches of two through the cluster kernel.  Heads against the oracle, the sampled
    indices against the host replica of the Philox stream (the utterance index of a later launch must be the global one)."""
    g = load_golden("dswnv_base")
    m = _model_from_golden(g)
    rng = np.random.default_rng(100 + B)
    Tf = 3
    U = g["kw"]["upsampling_factor"]
    N = Tf * U
    aux = rng.standard_normal((B, g["kw"]["n_aux"], Tf)).astype(np.float32)
    forced = rng.integers(0, 256, (B, N)).astype(np.int64)
    outs, heads = m.batch_fast_generate(torch.full((B, 1), 128, dtype=torch.long), torch.from_numpy(aux), [N] * B,
                                        forced=forced, seed=77, return_heads=True)
    info = m.engine().info
    assert info["kernel_variant"] == 2 and info["batch_tile"] == 2 and info["n_slots"] == -(-B // 2), info
    _head_close(heads, O.teacher_forced(g["cfg"], g["sd"], aux, forced), "chunked B=%d" % B, g)
    from shallow_wavenet_b200.engine import philox_uniform
    for b in range(B):
        u = philox_uniform(77, "dswnv", b, 0, N).reshape(N, 256)
        gum = heads[b] + _neg_log_q(u)
        assert (gum.argmax(-1) != outs[b]).mean() < 5e-3, b


def test_edge_cases_and_errors():
    from shallow_wavenet_b200._abi import SwnError
    g = load_golden("cswnv_seg5_lpc4_wav")
    m = _model_from_golden(g)
    aux = torch.from_numpy(g["aux"][:1, :, :1])                    # a single frame
    U = g["kw"]["upsampling_factor"]
    out = m.batch_fast_generate(torch.zeros(1, 5), aux, [U], seed=1)
    assert len(out) == 1 and out[0].shape == (U,)
    out = m.batch_fast_generate(torch.zeros(1, 5), aux, [U - 2], seed=1)   # not a multiple of seg: floor (cswnv:293)
    assert out[0].shape == (5,)
    with pytest.raises(SwnError, match="exceeds"):
        m.batch_fast_generate(torch.zeros(1, 5), aux, [U + 5], seed=1)
    with pytest.raises(ValueError):
        m.batch_fast_generate(torch.zeros(1, 5), aux, [U, U], seed=1)
    m.cpu()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.batch_fast_generate(torch.zeros(1, 5), aux, [U], seed=1)


def test_zero_noise_variate_and_index_validation():
    """Injected q == 0 for a class wins whatever its posterior - what the reference's argmax(p / q) does with it
    (p / 0 = inf); the Philox path never produces it (softmax_q floors q at 2^-24).  Out-of-range forced indices raise
    like the reference's OneHot scatter; a negative seed index follows Python's % (dswnv.py:77)."""
    g = load_golden("dswnv_base")
    m = _model_from_golden(g)
    B, N = g["samples"].shape
    aux = torch.from_numpy(g["aux"])
    seed_a = torch.full((B, 1), 128, dtype=torch.long)
    q = g["noise"].copy()
    _, heads = m.batch_fast_generate(seed_a, aux, [N] * B, noise=q, forced=g["samples"], return_heads=True)
    low = heads.argmin(-1)                                         # least likely class of every step
    for b in range(B):
        q[b, np.arange(N), low[b]] = 0.0
    for sel in (1, 2, 3):
        outs = m.batch_fast_generate(seed_a, aux, [N] * B, noise=q, forced=g["samples"], kernel_sel=sel)
        assert np.array_equal(np.stack(outs), low), sel
    assert np.array_equal(O.gumbel_argmax(heads.reshape(-1, 256), q.reshape(-1, 256)).reshape(B, N), low)
    bad = g["samples"].copy()
    bad[0, 3] = 256
    with pytest.raises(ValueError, match="forced class indices"):
        m.batch_fast_generate(seed_a, aux, [N] * B, noise=q, forced=bad)
    bad[0, 3] = -1
    with pytest.raises(ValueError, match="forced class indices"):
        m.batch_fast_generate(seed_a, aux, [N] * B, noise=q, forced=bad)
    a = m.batch_fast_generate(torch.full((B, 1), 128 - 256, dtype=torch.long), aux, [N] * B, noise=g["noise"], return_heads=True)[1]
    b_ = m.batch_fast_generate(torch.full((B, 1), 128 + 512, dtype=torch.long), aux, [N] * B, noise=g["noise"], return_heads=True)[1]
    assert np.array_equal(a, b_) and float(np.abs(a - g["heads"]).max()) < 1e-4


def test_edits_through_param_data_rebuild_engine():
    """p.data edits do not bump p._version: refresh_engine() after such an edit, or let the engine cache key
    carry a content checksum (check_weight_content)."""
    g = load_golden("dswnv_base")
    m = _model_from_golden(g)
    m.check_weight_content = True
    aux = torch.from_numpy(g["aux"])
    B, N = g["samples"].shape
    seed_a = torch.full((B, 1), 128, dtype=torch.long)
    m.batch_fast_generate(seed_a, aux, [N] * B, noise=g["noise"])
    m.out_2.bias.data.add_(1.0)
    _, h2 = m.batch_fast_generate(seed_a, aux, [N] * B, noise=g["noise"], forced=g["samples"], return_heads=True)
    assert np.allclose(h2 - 1.0, g["heads"], atol=1e-4)
    m.out_2.bias.data.sub_(1.0)
    m.refresh_engine()
    _, h3 = m.batch_fast_generate(seed_a, aux, [N] * B, noise=g["noise"], forced=g["samples"], return_heads=True)
    assert np.allclose(h3, g["heads"], atol=1e-4)


def test_weight_update_rebuilds_engine():
    g = load_golden("dswnv_base")
    m = _model_from_golden(g)
    aux = torch.from_numpy(g["aux"])
    B, N = g["samples"].shape
    seed_a = torch.full((B, 1), 128, dtype=torch.long)
    _, h1 = m.batch_fast_generate(seed_a, aux, [N] * B, noise=g["noise"], return_heads=True)
    with torch.no_grad():
        m.out_2.bias.add_(1.0)
    _, h2 = m.batch_fast_generate(seed_a, aux, [N] * B, noise=g["noise"], forced=g["samples"], return_heads=True)
    assert np.allclose(h2 - 1.0, g["heads"], atol=1e-4)


def test_host_buffer_entry_point():
    g = load_golden("dswnv_base")
    m = _model_from_golden(g)
    B, N = g["samples"].shape
    eng = m.engine()
    out = eng.generate_host(g["aux"], N, seed=3)
    ref = np.stack(m.batch_fast_generate(torch.full((B, 1), 128, dtype=torch.long), torch.from_numpy(g["aux"]), [N] * B, seed=3))
    assert out.dtype == np.int64 and np.array_equal(out, ref)


@pytest.mark.parametrize("name", ["dswnv_base", "dswnv_wavconv_k6", "cswnv_seg1_lpc0"])
def test_both_kernel_variants_agree(name):
    """Small batches with one sample per step run on the cluster kernel (variant 2) by default; the
    layer-group kernel (variant 1) must give the same answer."""
    g = load_golden(name)
    m = _model_from_golden(g)
    B, N = g["samples"].shape
    seed_a = _seed_audio(g["kind"], B, 1)
    res = {}
    for sel in (1, 2):
        outs, heads = m.batch_fast_generate(seed_a, torch.from_numpy(g["aux"]), [N] * B, noise=g["noise"],
                                            forced=g["samples"], return_heads=True, kernel_sel=sel)
        assert m.engine().info["kernel_variant"] == sel
        _head_close(heads, g["heads"], "%s variant %d" % (name, sel), g)
        res[sel] = (np.stack(outs), heads)
    if g["kind"] == "dswnv":
        assert np.array_equal(res[1][0], res[2][0])
    else:
        assert float(np.abs(res[1][0] - res[2][0]).max()) <= SAMPLE_ATOL
    assert float(np.abs(res[1][1] - res[2][1]).max()) <= 2e-5


SEG_FIXTURES = ["cswnv_seg2_lpc3", "cswnv_seg5_lpc0_conv2d", "cswnv_seg5_lpc4_h64", "cswnv_seg5_lpc4_wav"]


@pytest.mark.parametrize("name", SEG_FIXTURES)
def test_seg_cluster_kernel_matches_reference_fixture(name):
    """CSWNV with several samples per step on the cluster kernel (swn_coreseg.cuh, variant 2; one utterance per launch):
    teacher-forced heads and samples, then free running on the reference's own noise with per-utterance lengths."""
    g = load_golden(name)
    m = _model_from_golden(g)
    B, N = g["samples"].shape
    seed_a = _seed_audio(g["kind"], B, g["cfg"].eff_seg)
    outs, heads = m.batch_fast_generate(seed_a, torch.from_numpy(g["aux"]), [N] * B, noise=g["noise"],
                                        forced=g["samples"], return_heads=True, kernel_sel=2)
    info = m.engine().info
    assert info["kernel_variant"] == 2 and info["n_slots"] == B
    _head_close(heads, g["heads"], "%s seg cluster kernel" % name, g)
    assert float(np.abs(np.stack(outs) - g["samples"]).max()) <= SAMPLE_ATOL
    outs = m.batch_fast_generate(seed_a, torch.from_numpy(g["aux"]), g["n_samples_list"], noise=g["noise"], kernel_sel=2)
    assert [len(o) for o in outs] == g["n_samples_list"]
    for b, o in enumerate(outs):
        assert float(np.abs(o - g["samples"][b, :len(o)]).max()) <= 5e-5


@pytest.mark.parametrize("name,B,sel,side", [("cswnv_seg5_lpc4_h64", 3, 2, 3), ("cswnv_seg2_lpc3", 4, 2, 2),
                                             ("cswnv_seg5_lpc4_h64", 7, 0, 3), ("cswnv_seg5_lpc4_wav", 9, 0, 3),
                                             ("cswnv_seg5_lpc0_conv2d", 2, 0, 2), ("cswnv_seg2_lpc3", 5, 0, 3)])
def test_seg_cluster_kernel_ragged_lengths_run_each_utterance_for_its_own_length(name, B, sel, side):
    """The seg cluster kernel runs up to three utterances side by side (one core cluster each, shared helper CTAs), larger
    batches - up to nine by default, e.g. the recipe's decode batch of 7 - as launches of balanced sizes; with n_samples_list
    a launch runs for the longest of its utterances (the reference runs all to the longest of the batch and truncates,
    cswnv_shift1.py:291-295,426-428): each utterance equals the oracle's output truncated to its length, including lengths
    that are not multiples of seg and a 1-frame utterance."""
    g = load_golden(name)
    m = _model_from_golden(g)
    kw, cfg = g["kw"], g["cfg"]
    rng = np.random.default_rng(100 + B)
    Tmax, U, seg = 5, kw["upsampling_factor"], cfg.eff_seg
    frames = np.sort(rng.integers(1, Tmax + 1, B))
    frames[0], frames[-1] = 1, Tmax
    aux = np.zeros((B, kw["n_aux"], Tmax), np.float32)
    for b, t in enumerate(frames):
        aux[b, :, :t] = rng.standard_normal((kw["n_aux"], t))
    lens = [int(t) * U - (3 if b % 2 == 1 and b < B - 1 and t * U > 3 else 0) for b, t in enumerate(frames)]
    N = Tmax * U
    n_gen = (N // seg) * seg
    noise = rng.uniform(-0.4999, 0.5, (B, n_gen)).astype(np.float32)
    outs = m.batch_fast_generate(_seed_audio(g["kind"], B, seg), torch.from_numpy(aux), lens, noise=noise, kernel_sel=sel)
    info = m.engine().info
    assert info["kernel_variant"] == 2 and info["batch_tile"] == side and info["n_slots"] == B, info
    ref_smp = O.generate(cfg, g["sd"], aux, N, noise)
    for b, o in enumerate(outs):
        want = min(lens[b], n_gen)
        assert len(o) == want
        assert float(np.abs(o - ref_smp[b, :want]).max()) <= 5e-5


def test_seg_cluster_kernel_recipe_size_reference_fixture():
    """The recipe model (H=192, seg 5, lpc 4, k=7, L=6, R=690) on the seg cluster kernel against the unmodified reference's
    outputs (tests/golden_large, oracle/build_ref.py), and against the layer-group kernel on the same inputs."""
    g = load_large("large_cswnv_seg5_b2")
    m = g["model"].cuda().eval()
    B, N = g["samples"].shape
    seed_a = _seed_audio("cswnv", B, 5)
    res = {}
    for sel in (2, 1):
        outs, heads = m.batch_fast_generate(seed_a, torch.from_numpy(g["aux"]), [N] * B, noise=g["noise"],
                                            forced=g["samples"], return_heads=True, kernel_sel=sel)
        assert m.engine().info["kernel_variant"] == sel
        _head_close(heads, g["heads"], "large_cswnv_seg5_b2 v%d" % sel, g)
        assert float(np.abs(np.stack(outs) - g["samples"]).max()) <= SAMPLE_ATOL
        res[sel] = heads
    assert float(np.abs(res[1] - res[2]).max()) <= 5e-5
    outs = m.batch_fast_generate(seed_a, torch.from_numpy(g["aux"]), g["n_samples_list"], noise=g["noise"], kernel_sel=2)
    for b, o in enumerate(outs):
        assert float(np.abs(o - g["samples"][b, :len(o)]).max()) <= 5e-5


@pytest.mark.parametrize("kind,over,B", [("cswnv", dict(seg=1, lpc=4), 1), ("cswnv", dict(seg=1, lpc=4), 2), ("dswnv", dict(), 1),
                                         ("cswnv", dict(seg=5, lpc=4), 3), ("cswnv", dict(seg=1, lpc=4), 3),
                                         ("cswnv", dict(seg=1, lpc=4), 4), ("cswnv", dict(seg=1, lpc=4), 7), ("dswnv", dict(), 2)])
def test_cluster_kernels_with_late_helper_ctas_are_bit_identical(kind, over, B):
    """The helper CTAs of the cluster kernels normally finish a turn long before the core needs their sums, so the core's
    wait-for-a-late-helper path is almost never taken.  With every helper turn delayed (swn_gen_args::debug_helper_delay_ns)
    it is taken at every step: the per-step heads must not change by a bit.  (Before the reconvergence after the operand
    commit of swn_core2.cuh they were wrong by up to 0.7 at 20 us and irreproducible at 8 us.)"""
    from shallow_wavenet_b200 import synth
    kw = synth.cswnv_kwargs(22050) if kind == "cswnv" else synth.dswnv_kwargs(16000)
    kw.update(over)
    model = synth.make_model(kind, **kw).cuda()
    eng = model.engine()
    frames = 3
    n = frames * kw["upsampling_factor"]
    aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
    rng = np.random.default_rng(B)
    forced = (0.3 * rng.standard_normal((B, n))).astype(np.float32) if kind == "cswnv" else rng.integers(0, 256, (B, n))
    ref = None
    for delay in (0, 8000, 20000):
        eng.condition(aux)
        # (B >= 3 with one sample per step: automatic choice = two core clusters of two utterances side by side per launch)
        out, heads = eng.generate(n, seed=7, kernel_sel=2 if B == 1 else 0, forced=forced, return_heads=True, helper_delay_ns=delay)
        assert eng.info["kernel_variant"] == 2
        if B > 3:
            assert eng.info["batch_tile"] == 2 and eng.info["n_slots"] == -(-B // 2), eng.info
        elif B == 2 and kind == "cswnv" and over.get("seg") == 1:       # one core cluster per utterance
            assert eng.info["batch_tile"] == 1 and eng.info["n_slots"] == 2, eng.info
        h = heads.cpu().numpy()
        if ref is None:
            ref = h
        else:
            assert np.array_equal(h, ref), "helper delay %d ns: max |diff| %.3e" % (delay, float(np.abs(h - ref).max()))


def test_cluster_kernel_is_used_for_the_headline_workload():
    model, cfg, sd, aux, kw = _recipe_case("dswnv", 16000, 1, 2, 3)
    N = 2 * kw["upsampling_factor"]
    model.batch_fast_generate(torch.full((1, 1), 128, dtype=torch.long), torch.from_numpy(aux), [N], seed=1)
    info = model.engine().info
    assert info["kernel_variant"] == 2 and info["grid"] % 16 == 0 and info["n_slots"] == 1


# ----------------------------------------------------------------------------- full-size properties
def test_full_size_cfg2_tail_and_sampler_properties():
    """BASELINE.json configs[1]: DSWNV, 16 kHz, batch 1, 5 s = 80000 samples, free running with Philox
    noise.  Checked: (1) every sampled index equals the argmax recomputed on the host from the kernel's
    own logits and the replicated Philox stream; (2) the heads of the last steps equal the oracle
    evaluated over the last receptive field of the kernel's own output (finite receptive field)."""
    from shallow_wavenet_b200 import engine as E
    model, cfg, sd, aux, kw = _recipe_case("dswnv", 16000, 1, 1000, 1)
    N = 1000 * kw["upsampling_factor"]
    outs, heads = model.batch_fast_generate(torch.full((1, 1), 128, dtype=torch.long), torch.from_numpy(aux), [N],
                                            seed=11, return_heads=True)
    smp = outs[0]
    assert smp.shape == (N,) and smp.min() >= 0 and smp.max() <= 255
    assert len(np.unique(smp)) > 32                                # not stuck
    chunk = 8000
    mism = 0
    for n0 in range(0, N, chunk):
        u = E.philox_uniform(11, "dswnv", 0, n0, chunk)
        g_ = heads[0, n0:n0 + chunk] + _neg_log_q(u)
        mism += int((np.argmax(g_, -1) != smp[n0:n0 + chunk]).sum())
    assert mism <= N * 1e-3, mism                                  # host log vs device logf near-ties only
    n_tail = 16
    ref_tail = O.teacher_forced_tail(cfg, sd, aux, smp[None, :], n_tail)
    _head_close(heads[:, -n_tail:], ref_tail, "cfg2 tail", kind="dswnv", kw=kw)


def test_full_size_cfg1_on_gpu_tail_property():
    """BASELINE.json configs[0] shape (CSWNV seg=5 lpc=4, 22.05 kHz, 1 s) on the GPU."""
    model, cfg, sd, aux, kw = _recipe_case("cswnv", 22050, 1, 200, 1)
    N = 200 * kw["upsampling_factor"]
    outs, heads = model.batch_fast_generate(torch.zeros(1, 5), torch.from_numpy(aux), [N], seed=12, return_heads=True)
    smp = outs[0]
    assert smp.shape == (N,) and np.abs(smp).max() <= 1.0 and np.isfinite(smp).all()
    n_tail = 8
    ref_tail = O.teacher_forced_tail(cfg, sd, aux, smp[None, :], n_tail)
    _head_close(heads[:, -n_tail:], ref_tail, "cfg1 tail", kind="cswnv", kw=kw)


# ----------------------------------------------------------------------------- variant 3: pipelined batch kernel
_V3_FIXTURES = [n for n in golden_names() if n != "dswnv_audio_in"]      # audio_in_flag stays on variant 1


@pytest.mark.parametrize("engine", [3, 4])          # 3: tensor-core cells where they apply, 4: fp32 FFMA cells
@pytest.mark.parametrize("tile", [1, 2])
@pytest.mark.parametrize("name", _V3_FIXTURES)
def test_batch_kernel_teacher_forced_matches_reference_fixture(name, tile, engine):
    g = load_golden(name)
    m = _model_from_golden(g)
    B, N = g["samples"].shape
    outs, heads = m.batch_fast_generate(_seed_audio(g["kind"], B, g["cfg"].eff_seg), torch.from_numpy(g["aux"]),
                                        [N] * B, noise=g["noise"], forced=g["samples"], return_heads=True,
                                        kernel_sel=engine, batch_tile=tile)
    info = m.engine().info
    assert info["kernel_variant"] == engine and info["batch_tile"] == tile and info["n_slots"] == -(-B // tile)
    _head_close(heads, g["heads"], "%s v%d tile %d" % (name, engine, tile), g, tol=_tol(m))
    smp = np.stack(outs)
    if g["kind"] == "dswnv":
        assert np.array_equal(smp, g["samples"])
    else:
        assert float(np.abs(smp - g["samples"]).max()) <= SAMPLE_ATOL


@pytest.mark.parametrize("name", _V3_FIXTURES)
def test_batch_kernel_free_running_matches_reference_fixture(name):
    g = load_golden(name)
    m = _model_from_golden(g)
    B, N = g["samples"].shape
    outs = m.batch_fast_generate(_seed_audio(g["kind"], B, g["cfg"].eff_seg), torch.from_numpy(g["aux"]),
                                 g["n_samples_list"], noise=g["noise"], kernel_sel=3)
    assert m.engine().info["kernel_variant"] == 3          # every fixture is small enough for the tensor-core cells
    for b, o in enumerate(outs):
        ref = g["samples"][b, :len(o)]
        if g["kind"] == "dswnv":
            assert np.array_equal(o, ref)
        else:
            assert float(np.abs(o - ref).max()) <= 5e-5


@pytest.mark.parametrize("name,B,tile", [("dswnv_base", 11, 4), ("dswnv_base", 37, 8), ("dswnv_wavconv_k6", 19, 16),
                                         ("cswnv_seg5_lpc4_wav", 13, 4), ("cswnv_seg5_lpc4_h64", 21, 8),
                                         ("cswnv_seg2_lpc3", 9, 2), ("cswnv_seg1_lpc0", 33, 0)])
def test_batch_kernel_many_slots_free_running_vs_oracle(name, B, tile):
    """More utterances than one slot holds: the ring pipeline (slots in flight through the layer groups,
    several head CTAs) against the oracle, free running on injected noise."""
    g = load_golden(name)
    m = _model_from_golden(g)
    kw, cfg = g["kw"], g["cfg"]
    rng = np.random.default_rng(B)
    Tf, U = 4, kw["upsampling_factor"]
    N = (Tf * U // cfg.eff_seg) * cfg.eff_seg
    aux = rng.standard_normal((B, kw["n_aux"], Tf)).astype(np.float32)
    if g["kind"] == "dswnv":
        noise = rng.exponential(1.0, (B, N, 256)).astype(np.float32)
    else:
        noise = rng.uniform(-0.4999, 0.5, (B, N)).astype(np.float32)
    outs, heads = m.batch_fast_generate(_seed_audio(g["kind"], B, cfg.eff_seg), torch.from_numpy(aux), [N] * B,
                                        noise=noise, return_heads=True, batch_tile=tile)
    info = m.engine().info
    assert info["kernel_variant"] in (3, 4) and info["n_slots"] == -(-B // info["batch_tile"])
    ref_smp, ref_heads = O.generate(cfg, g["sd"], aux, N, noise, return_heads=True)
    _head_close(heads, ref_heads, "%s B=%d" % (name, B), g, tol=_tol(m))
    if g["kind"] == "dswnv":
        assert np.array_equal(np.stack(outs), ref_smp)
    else:
        assert float(np.abs(np.stack(outs) - ref_smp).max()) <= 5e-5


@pytest.mark.parametrize("kind,fs,B,tile,over,variant", [("cswnv", 22050, 9, 4, dict(seg=5, lpc=4), 3),
                                                          ("cswnv", 22050, 17, 8, dict(seg=5, lpc=4), 3),
                                                          ("cswnv", 22050, 17, 8, dict(seg=5, lpc=4), 4),
                                                          ("cswnv", 22050, 35, 16, dict(seg=5, lpc=4), 3),
                                                          ("cswnv", 22050, 6, 2, dict(seg=10, lpc=4), 3),
                                                          ("cswnv", 22050, 12, 4, dict(seg=1, lpc=4), 3),
                                                          ("dswnv", 24000, 20, 8, dict(), 4),
                                                          ("dswnv", 16000, 33, 8, dict(), 4)])
def test_batch_kernel_recipe_size_teacher_forced_vs_oracle(kind, fs, B, tile, over, variant):
    """Recipe sizes: CSWNV (H=192) runs its cells on the tensor cores (variant 3: tcgen05, three tf32 passes,
    weights in tensor memory); DSWNV (H=256) does not fit tensor memory and stays on the fp32 FFMA cells (4)."""
    model, cfg, sd, aux, kw = _recipe_case(kind, fs, B, 2, 21, **over)
    seg = cfg.eff_seg
    N = 2 * kw["upsampling_factor"]
    rng = np.random.default_rng(8)
    if kind == "cswnv":
        forced = np.clip(0.3 * rng.standard_normal((B, N)), -1, 1).astype(np.float32)
        noise = rng.uniform(-0.4999, 0.5, (B, N)).astype(np.float32)
    else:
        forced = rng.integers(0, 256, (B, N)).astype(np.int64)
        noise = rng.exponential(1.0, (B, N, 256)).astype(np.float32)
    outs, heads = model.batch_fast_generate(_seed_audio(kind, B, seg), torch.from_numpy(aux), [N] * B, noise=noise,
                                            forced=forced, return_heads=True, batch_tile=tile, kernel_sel=variant)
    info = model.engine().info
    assert info["kernel_variant"] == variant and info["batch_tile"] == tile
    ref_heads = O.teacher_forced(cfg, sd, aux, forced)
    _head_close(heads, ref_heads, "%s recipe B=%d v%d" % (kind, B, variant), kind=kind, kw=kw, tol=_tol(model))
    if kind == "dswnv":
        mine = O.gumbel_argmax(heads.reshape(-1, 256), noise.reshape(-1, 256)).reshape(B, N)
        assert np.array_equal(np.stack(outs), mine)


def _tail_check(cfg, sd, aux, smp, heads, picks, frames_tail, U, n_tail, what, kind, kw, tol):
    """Finite receptive field: the heads of the last steps only depend on the last frames of aux and the last
    samples, so the oracle is evaluated on that window (seconds instead of the reference's O(N^2) loop)."""
    for b in picks:
        ref_tail = O.teacher_forced_tail(cfg, sd, aux[b:b + 1, :, -frames_tail:], smp[b:b + 1, -frames_tail * U:], n_tail)
        _head_close(heads[b:b + 1, -n_tail:], ref_tail, "%s utterance %d tail" % (what, b), kind=kind, kw=kw, tol=tol)


def test_full_size_cfg3_batch64_properties():
    """BASELINE.json configs[2]: CSWNV seg=5 lpc=4, 22.05 kHz, batch 64 x 5 s (110000 samples each), free
    running on Philox noise through the pipelined batch kernel with tensor-core cells."""
    B, frames = 64, 1000
    model, cfg, sd, aux, kw = _recipe_case("cswnv", 22050, B, frames, 3)
    U = kw["upsampling_factor"]
    N = frames * U
    outs, heads = model.batch_fast_generate(torch.zeros(B, 5), torch.from_numpy(aux), [N] * B, seed=21, return_heads=True)
    info = model.engine().info
    assert info["kernel_variant"] == 3 and info["n_slots"] * info["batch_tile"] >= B
    smp = np.stack(outs)
    assert smp.shape == (B, N) and np.isfinite(smp).all() and np.abs(smp).max() <= 1.0
    assert np.isfinite(heads).all()
    assert len({smp[b, -2000:].tobytes() for b in range(B)}) == B          # every utterance is its own stream
    assert float(np.abs(smp[:, -20000:]).mean()) > 1e-4                     # not silent
    _tail_check(cfg, sd, aux, smp, heads, (0, 37, 63), 20, U, 6, "cfg3", "cswnv", kw, _tol(model))


def test_cfg5_shape_batch128_properties():
    """BASELINE.json configs[4] shape per GPU (DSWNV, 24 kHz, 128 utterances), 2 s instead of 60 s: indices are the
    arg-max recomputed from the kernel's own logits and the replicated Philox stream; tail heads equal the oracle's."""
    from shallow_wavenet_b200 import engine as E
    B, frames = 128, 400
    model, cfg, sd, aux, kw = _recipe_case("dswnv", 24000, B, frames, 4)
    U = kw["upsampling_factor"]
    N = frames * U
    outs, heads = model.batch_fast_generate(torch.full((B, 1), 128, dtype=torch.long), torch.from_numpy(aux), [N] * B,
                                            seed=31, return_heads=True)
    info = model.engine().info
    assert info["kernel_variant"] in (3, 4) and info["n_slots"] > 1
    smp = np.stack(outs)
    assert smp.shape == (B, N) and smp.min() >= 0 and smp.max() <= 255
    for b in (0, 77, 127):
        n0 = N - 4000
        u = E.philox_uniform(31, "dswnv", b, n0, 4000)
        g_ = heads[b, n0:] + _neg_log_q(u)
        assert int((np.argmax(g_, -1) != smp[b, n0:]).sum()) <= 8           # host log vs device logf near-ties only
    _tail_check(cfg, sd, aux, smp, heads, (0, 127), 20, U, 8, "cfg5 shape", "dswnv", kw, _tol(model))


@pytest.mark.parametrize("engine", [3, 4])
def test_batch_kernel_cell0_builds_its_input_option(engine, monkeypatch):
    """SWN_BUILD_X=1: the CTAs of cell 0 evaluate the causal input stage themselves from the samples the tail
    publishes (kept as an option: measured slower than the tail building the tile); same results."""
    monkeypatch.setenv("SWN_BUILD_X", "1")
    g = load_golden("cswnv_seg5_lpc4_h64")
    m = _model_from_golden(g)
    rng = np.random.default_rng(17)
    B, Tf, U = 11, 4, g["kw"]["upsampling_factor"]
    N = Tf * U
    aux = rng.standard_normal((B, g["kw"]["n_aux"], Tf)).astype(np.float32)
    eps = rng.uniform(-0.4999, 0.5, (B, N)).astype(np.float32)
    outs, heads = m.batch_fast_generate(torch.zeros(B, 5), torch.from_numpy(aux), [N] * B, noise=eps, return_heads=True,
                                        kernel_sel=engine, batch_tile=4)
    assert m.engine().info["kernel_variant"] == engine
    ref_smp, ref_heads = O.generate(g["cfg"], g["sd"], aux, N, eps, return_heads=True)
    _head_close(heads, ref_heads, "build_x engine %d" % engine, g, tol=_tol(m))
    assert float(np.abs(np.stack(outs) - ref_smp).max()) <= 5e-5


@pytest.mark.parametrize("name,B,tile,engine", [("cswnv_seg5_lpc4_h64", 21, 4, 3), ("cswnv_seg5_lpc4_h64", 21, 4, 4),
                                                ("dswnv_base", 19, 2, 3), ("cswnv_seg2_lpc3", 7, 1, 3)])
def test_batch_kernel_ragged_lengths_leave_the_ring_early(name, B, tile, engine):
    """Utterances of unequal length (decode_generator zero-pads the aux and passes n_samples_list,
    decode_cswnv_laplace-shift1.py:99-111): a pipeline slot stops once its longest utterance is complete; every
    utterance still equals the oracle's output truncated to its own length (cswnv_shift1.py:426-428)."""
    g = load_golden(name)
    m = _model_from_golden(g)
    kw, cfg = g["kw"], g["cfg"]
    rng = np.random.default_rng(B + tile)
    Tmax, U, seg = 6, kw["upsampling_factor"], cfg.eff_seg
    frames = np.sort(rng.integers(1, Tmax + 1, B))            # the driver sorts by length
    frames[-1] = Tmax
    aux = np.zeros((B, kw["n_aux"], Tmax), np.float32)
    for b, t in enumerate(frames):
        aux[b, :, :t] = rng.standard_normal((kw["n_aux"], t))
    lens = [int(t) * U - (3 if b % 4 == 1 and t * U > 3 else 0) for b, t in enumerate(frames)]   # some not multiples of seg
    N = Tmax * U
    n_gen = (N // seg) * seg
    if g["kind"] == "dswnv":
        noise = rng.exponential(1.0, (B, n_gen, 256)).astype(np.float32)
    else:
        noise = rng.uniform(-0.4999, 0.5, (B, n_gen)).astype(np.float32)
    outs = m.batch_fast_generate(_seed_audio(g["kind"], B, seg), torch.from_numpy(aux), lens, noise=noise,
                                 batch_tile=tile, kernel_sel=engine)
    info = m.engine().info
    assert info["kernel_variant"] == engine and info["n_slots"] == -(-B // tile)
    ref_smp = O.generate(cfg, g["sd"], aux, N, noise)
    for b, o in enumerate(outs):
        want = min(lens[b], n_gen)
        assert len(o) == want
        if g["kind"] == "dswnv":
            assert np.array_equal(o, ref_smp[b, :want])
        else:
            assert float(np.abs(o - ref_smp[b, :want]).max()) <= 5e-5


@pytest.mark.parametrize("engine", [3, 4])
def test_batch_kernel_throughput_options_forced_on_small_rings(engine, monkeypatch):
    """The prefetch team of the gated cells (SWN_P1_SPLIT) and the two half-size tail instances (SWN_TAIL_INST=2) are
    switched on by the plan from 16 slots up; forced here on small, ragged rings (incl. the case where a single slot
    remains in the ring) so that their hand-offs are exercised by the regular suite."""
    monkeypatch.setenv("SWN_P1_SPLIT", "1")
    monkeypatch.setenv("SWN_TAIL_INST", "2")
    test_batch_kernel_ragged_lengths_leave_the_ring_early("cswnv_seg5_lpc4_h64", 21, 4, engine)
    test_batch_kernel_ragged_lengths_leave_the_ring_early("cswnv_seg2_lpc3", 7, 1, engine)
    test_batch_kernel_ragged_lengths_leave_the_ring_early("dswnv_base", 19, 2, engine)
