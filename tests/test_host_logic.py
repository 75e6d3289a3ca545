"""CPU tests: boundary classes, ABI library symbols, sharding logic (gloo, world_size 2)."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from _util import ROOT, golden_names, load_golden
import shallow_wavenet_b200 as S
from shallow_wavenet_b200 import _abi, sharding


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "swn_b200.h")).read()
    declared = set(re.findall(r"\b(swn_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s
    _abi.load()


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _abi.load()
    cfg = _abi.SwnConfig()
    h = ctypes.c_void_p()
    assert lib.swn_create(ctypes.byref(cfg), 0, ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.swn_last_error() or b"CPU fallback" in lib.swn_last_error()
    m = S.DSWNV(n_aux=4, hid_chn=16, skip_chn=16, dilation_depth=2, dilation_repeat=1, kernel_size=3, upsampling_factor=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.batch_fast_generate(torch.full((1, 1), 128), torch.zeros(1, 4, 2), [8])


@pytest.mark.parametrize("name", golden_names())
def test_state_dict_layout_matches_reference(name):
    g = load_golden(name)
    cls = S.CSWNV if g["kind"] == "cswnv" else S.DSWNV
    m = cls(**g["kw"])
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["sd"].keys())            # same names, same order
    for k, v in g["sd"].items():
        assert tuple(sd[k].shape) == v.shape, k
    m.load_state_dict({k: torch.from_numpy(v) for k, v in g["sd"].items()}, strict=True)


def test_recipe_constructor_attributes():
    m = S.CSWNV(wav_conv_flag=True)                            # class defaults = recipe (run.sh:164-230)
    assert m.receptive_field == 690 and m.padding == [6, 42, 294, 6, 42, 294]
    assert m.dil_facts == [0, 1, 2, 0, 1, 2] and m.seg == 5 and m.lpc == 4 and m.lpc_offset == 1
    assert m.in_aux_dim == 54 * 9 * 5
    assert sum(p.numel() for p in m.parameters()) == 9590535
    d = S.DSWNV(hid_chn=256, dilation_repeat=2, kernel_size=7)
    assert sum(p.numel() for p in d.parameters()) == 8255673
    assert d.receptive_field == 690 and d.in_tot_dim == 486
    d2 = S.DSWNV()                                             # class defaults: k=6, repeat 3
    assert d2.receptive_field == sum(d2.padding) + 5 and len(d2.dil_facts) == 9


def test_initialize_matches_reference_semantics():
    torch.manual_seed(0)
    m = S.CSWNV(n_aux=4, hid_chn=16, skip_chn=16, upsampling_factor=6, seg=2, lpc=2)
    m.apply(S.initialize)
    assert float(m.upsampling.conv.weight.min()) == 1.0 and float(m.upsampling.conv.bias.abs().max()) == 0.0
    w = m.dil_h[0].conv.weight
    bound = (6.0 / ((16 + 32) * 3)) ** 0.5                     # xavier_uniform on [32,16,3]
    assert float(w.abs().max()) <= bound + 1e-6 and float(w.std()) > 0.3 * bound


def test_mu_law_matches_oracle():
    from oracle import swn_oracle as O
    x = np.linspace(-1, 1, 1001)
    assert np.array_equal(S.encode_mu_law(x), O.encode_mu_law(x))
    y = np.arange(256)
    assert np.array_equal(S.decode_mu_law(y), O.decode_mu_law(y))


def test_sharding_rules():
    items = ["u%02d" % i for i in range(11)]
    parts = sharding.split_for_gpus(items, 4)
    assert [len(p) for p in parts] == [3, 3, 3, 2] and sum(parts, []) == items
    lengths = [5, 3, 9, 3, 7, 1, 8]
    batches = sharding.decode_batches(lengths, 3)
    assert len(batches) == 3 and sorted(sum(batches, [])) == list(range(7))
    flat = [lengths[i] for b in batches for i in b]
    assert flat == sorted(lengths)
    aux = sharding.pad_aux([np.ones((2, 3)), 2 * np.ones((4, 3))])
    assert aux.shape == (2, 3, 4) and aux[0, :, 2:].sum() == 0 and aux[1].min() == 2


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = list(range(13))
    mine = sharding.rank_shard(items, rank, world)
    thr = sharding.aggregate_throughput(100.0 * len(mine), 1.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mine, thr, gathered))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res[0][1] == list(range(7)) and res[1][1] == list(range(7, 13))
    assert sorted(sum(res[0][3], [])) == list(range(13))       # every utterance exactly once
    for r in res:
        assert abs(r[2] - 1300.0 / 2.0) < 1e-9                 # sum of samples / max of time
