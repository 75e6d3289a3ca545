#!/usr/bin/env python
"""Throughput of the time-parallel forward (swn_forward) on a B200: time steps/s and algorithmic TFLOP/s against the
tensor roofline (MEASURED_PEAKS.json bf16 sustained; same-precision ceiling = peak / 6 for three tf32 passes).
usage: python tools/gpu_forward_bench.py [cswnv|dswnv] [B] [frames] [reps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shallow_wavenet_b200 import synth  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cswnv"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
fs = 22050 if kind == "cswnv" else 16000
kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
model = synth.make_model(kind, **kw).cuda()
H, S, K, L = kw["hid_chn"], kw["skip_chn"], kw["kernel_size"], kw["dilation_depth"] * kw["dilation_repeat"]
seg = kw.get("seg", 1) if kind == "cswnv" else 1
n1 = S if kind == "cswnv" else 256
n2 = 2 * seg + kw.get("lpc", 0) if kind == "cswnv" else 256
macs = L * (2 * H * H * K + H * S) + S * n1 + n1 * n2          # per time step, every multiply-add of the reference once
aux = torch.from_numpy(synth.synth_aux(kw["n_aux"], frames, B)).cuda()
Ta = frames * kw["upsampling_factor"] - seg
rng = np.random.default_rng(0)
if kind == "cswnv":
    audio = torch.from_numpy(np.clip(0.3 * rng.standard_normal((B, Ta)), -1, 1).astype(np.float32)).cuda()
else:
    audio = torch.from_numpy(rng.integers(0, 256, (B, Ta))).cuda()
eng = model.engine()
ms = []
for i in range(reps + 2):
    eng.condition(aux)
    eng.forward(audio)
    if i >= 2:
        ms.append(eng.info["last_gen_ms"])
T = Ta - seg + 1
steps = B * T
t = float(np.mean(ms)) * 1e-3
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = peaks.get("bf16_tflops_sustained") or 1400.0
tf = 2.0 * macs * steps / t / 1e12
print(json.dumps({"what": "swn_forward (%s recipe size), B=%d x %d time steps" % (kind.upper(), B, T), "ms": t * 1e3, "ms_each": ms,
                  "time_steps_per_s": steps / t, "x_real_time": steps / t / fs, "algorithmic_flop_per_time_step": 2.0 * macs,
                  "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                               "peak_tf32x3_equivalent": peak / 6.0, "frac_of_tf32x3_equivalent": tf / (peak / 6.0)},
                  "kernel_launches_per_forward": 3 * L + 6}))
