"""Throughput of the noise-shaping restore filter on the GPU next to the oracle on the host (bounded sample)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import shallow_wavenet_b200.noise_shaping as NS
from oracle import noise_shaping_oracle as NO

fs, alpha = 22050, 0.455
rng = np.random.default_rng(0)
mean = rng.standard_normal(55); mean[5:] = 0.6 * rng.standard_normal(50) / (1 + 0.25 * np.arange(50)); mean[5] = -4.0
res = {"what": "noise-shaping restore filter (MLSA order 49, pd 4, alpha %.3f + 255-tap low cut), float64, fs %d" % (alpha, fs)}
for B, secs in ((1, 5.0), (64, 5.0), (512, 10.0)):
    n = int(secs * fs)
    wavs = [np.clip(0.3 * rng.standard_normal(n), -1, 1) for _ in range(B)]
    NS.noise_shaping(wavs[:1], mean, fs=fs, mcep_alpha=alpha)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); out = NS.noise_shaping(wavs, mean, fs=fs, mcep_alpha=alpha); dt = time.perf_counter() - t0
    res["B=%d x %.0f s" % (B, secs)] = {"ms_end_to_end_host_buffers": dt * 1e3, "samples_per_s": B * n / dt, "x_real_time": B * n / dt / fs}
x = wavs[0][:4410]
t0 = time.perf_counter(); ref = NO.noise_shaping(x, mean, fs, 5, alpha); dt = time.perf_counter() - t0
res["oracle (pure Python restatement) on the host, 0.2 s of audio"] = {"samples_per_s": len(x) / dt}
res["max_abs_diff_vs_oracle_on_that_slice"] = float(np.abs(NS.noise_shaping([x], mean, fs=fs, mcep_alpha=alpha)[0] - ref).max())
print(json.dumps(res))
