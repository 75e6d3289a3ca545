# -*- coding: utf-8 -*-
"""Utterance sharding of the decode drivers (host logic, no collective).

The reference splits the utterance list into ``n_gpus`` contiguous chunks and forks one
process per GPU (src/bin/decode_cswnv_laplace-shift1.py:200-201,262-274); inside a process it
sorts by length and ``np.array_split``s into ceil(n/batch_size) batches (:77-84).  The same
rules, as pure functions usable with one process per GPU under torch.distributed."""
import math

import numpy as np


def split_for_gpus(items, n_gpus):
    """decode_cswnv_laplace-shift1.py:200-201: contiguous np.array_split over the GPUs."""
    return [list(c) for c in np.array_split(np.asarray(list(items), dtype=object), n_gpus)]


def rank_shard(items, rank, world_size):
    return split_for_gpus(items, world_size)[rank]


def decode_batches(lengths, batch_size):
    """decode_generator (:77-84): indices sorted by length with numpy's DEFAULT argsort - the order of equal
    lengths is whatever the reference gets -, then split into ceil(n / batch_size) nearly equal batches."""
    idx = np.argsort(np.asarray(lengths))
    n_batch = int(math.ceil(len(idx) / float(batch_size)))
    return [b.tolist() for b in np.array_split(idx, n_batch)] if n_batch else []


def pad_aux(aux_list):
    """pad_list (:30-48) + transpose: list of [T_i, n_aux] -> zero padded [B, n_aux, T_max] float32."""
    T = max(a.shape[0] for a in aux_list)
    out = np.zeros((len(aux_list), aux_list[0].shape[1], T), np.float32)
    for i, a in enumerate(aux_list):
        out[i, :, :a.shape[0]] = np.asarray(a, np.float32).T
    return out


def aggregate_throughput(samples_local, seconds_local, group=None):
    """Whole-job samples/s: sum of samples over ranks / max of time over ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(samples_local)], dtype=torch.float64)
    s = torch.tensor([float(seconds_local)], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend(group) == "nccl":
            t, s = t.cuda(), s.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(s, op=dist.ReduceOp.MAX, group=group)
    return float(t.item()) / float(s.item())
