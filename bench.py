#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Benchmark of the B200-native shallow-wavenet generation path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg1|cfg3|...]

One "step" = one pass of the hot path over one batch of synthetic utterances: conditioning +
warm-up + the whole autoregressive loop (``batch_fast_generate``).  The default workload is
BASELINE.json configs[1] (DSWNV mu-law softmax, 16 kHz, batch 1, 5 s); with ``--gpus N`` every
rank generates its own copy of that workload (utterances shard with no collective: weak
scaling).  Rank 0 prints ONE JSON line; ``value`` / ``e2e`` / ``roofline`` are
those of the headline workload, and the same line carries two sub-records measured in the same
run: ``batched.cfg3`` (configs[2]: 64 x 5 s on the tensor-core batch kernel, 2 steps) and, under
``--gpus N > 1`` (or ``--sharded-at-n1``), ``sharded.cfg4`` (configs[3]: 512 utterances of
unequal length split over the ranks with ``sharding.split_for_gpus``, STRONG scaling).

``--impl reference`` times the UNMODIFIED reference (``oracle/_ref``: a verbatim, git-ignored
copy of its two ``src/nets`` modules made by ``__graft_entry__.build()``; it travels to the GPU
box with the snapshot) on the host cores, on a bounded sample of the same workload; only when
that copy is missing does it fall back to the oracle port, and says so (``kind: "port"``).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, fs, ctor overrides, B per GPU, frames, description)
    "cfg1": ("cswnv", 22050, dict(seg=5, lpc=4), 1, 200,
             "cfg1: CSWNV Laplace seg=5 lpc=4 H=192 S=256 k=7 L=6, 22.05 kHz, batch 1, 1 s (22000 samples)"),
    "cfg1x3": ("cswnv", 22050, dict(seg=5, lpc=4), 3, 200,
               "cfg1 x 3: the cfg1 model, 3 utterances of 1 s side by side (one core cluster each, shared helper CTAs)"),
    "cfg1x7": ("cswnv", 22050, dict(seg=5, lpc=4), 7, 200,
               "cfg1 x 7: the cfg1 model at the recipe's decode batch (decode_batch_size=7, egs/shallow-wavenet/run.sh:280), 7 x 1 s"),
    "seg1x7": ("cswnv", 22050, dict(seg=1, lpc=4), 7, 200,
               "recipe default x 7: CSWNV Laplace seg=1 lpc=4 H=192 (model=laplace, seg=1: egs/shallow-wavenet/run.sh:159-219) at the "
               "recipe's decode batch of 7, 7 x 1 s"),
    "cfg2x7": ("dswnv", 16000, dict(), 7, 200,
               "cfg2 x 7: the cfg2 model (DSWNV H=256) at the recipe's decode batch of 7, 7 x 1 s (launches of two: its helper taps "
               "leave no room for a second core cluster)"),
    "cfg2": ("dswnv", 16000, dict(), 1, 1000,
             "cfg2: DSWNV mu-law softmax H=256 S=256 Q=256 k=7 L=6, 16 kHz, batch 1, 5 s (80000 samples)"),
    "cfg3": ("cswnv", 22050, dict(seg=5, lpc=4), 64, 1000,
             "cfg3: CSWNV Laplace seg=5 lpc=4 H=192, 22.05 kHz, batch 64 x 5 s (110000 samples each)"),
    "cfg4": ("cswnv", 22050, dict(seg=5, lpc=4), 64, 2000,
             "cfg4: CSWNV Laplace seg=5 lpc=4 H=192, 22.05 kHz, batch 64 per GPU x 10 s (the 8-GPU shard of batch 512)"),
    "cfg4_256": ("cswnv", 22050, dict(seg=5, lpc=4), 256, 2000,
                 "cfg4: CSWNV Laplace seg=5 lpc=4 H=192, 22.05 kHz, batch 256 per GPU x 10 s (the 2-GPU shard of batch 512)"),
    "cfg3s": ("cswnv", 22050, dict(seg=5, lpc=4), 64, 40,
              "cfg3 short: CSWNV seg=5 lpc=4 H=192, 22.05 kHz, batch 64 x 0.2 s (4400 samples each)"),
    "cfg5s": ("dswnv", 24000, dict(), 128, 40,
              "cfg5 short: DSWNV H=256, 24 kHz, batch 128 x 0.2 s (4800 samples each)"),
    "cfg5": ("dswnv", 24000, dict(), 128, 12000,
             "cfg5: DSWNV mu-law softmax H=256, 24 kHz, batch 128 per GPU x 60 s"),
    "cfg2s": ("dswnv", 16000, dict(), 1, 100,
              "cfg2 short: DSWNV H=256, 16 kHz, batch 1, 0.5 s (8000 samples)"),
}


def model_kwargs(kind, fs, over):
    from shallow_wavenet_b200 import synth
    kw = synth.cswnv_kwargs(fs) if kind == "cswnv" else synth.dswnv_kwargs(fs)
    kw.update(over)
    return kw


def in_loop_macs(kind, kw):
    """Algorithmic in-loop MACs per generated sample and in-loop weight count."""
    H, S, k = kw["hid_chn"], kw["skip_chn"], kw["kernel_size"]
    L = kw["dilation_depth"] * kw["dilation_repeat"]
    if kind == "cswnv":
        seg, lpc = kw["seg"], kw["lpc"]
        macs = H + H * H * k + L * (2 * H * H * k + H * S) + (S * S + S * (2 * seg + lpc)) / float(seg)
        weights = H * H * k + L * (2 * H * H * k + H * S) + S * S + S * (2 * seg + lpc)
    else:
        Q = kw["n_quantize"]
        macs = k * H + L * (2 * H * H * k + H * S) + S * Q + Q * Q
        weights = H * Q * k + L * (2 * H * H * k + H * S) + S * Q + Q * Q
    return macs, weights


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super(ClockSampler, self).__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _cpu_inputs(kind, kw, frames, sample_frames):
    from shallow_wavenet_b200 import synth
    model = synth.make_model(kind, **kw)
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    aux = synth.synth_aux(kw["n_aux"], frames, 1)[:, :, :sample_frames]
    return sd, np.ascontiguousarray(aux), sample_frames * kw["upsampling_factor"]


def cpu_port_run(kind, kw, frames, sample_frames, threads):
    """Time the CPU port of the reference algorithm (oracle) on a bounded sample of the workload."""
    import torch
    from oracle import swn_oracle as O
    torch.set_num_threads(threads)
    sd, aux, n = _cpu_inputs(kind, kw, frames, sample_frames)
    cfg = O.SWNConfig(kind=kind, **dict(kw))
    rng = np.random.default_rng(0)
    if kind == "cswnv":
        noise = rng.uniform(-0.4999, 0.5, (1, n)).astype(np.float32)
    else:
        noise = rng.exponential(1.0, (1, n, 256)).astype(np.float32)
    t0 = time.perf_counter()
    O.generate(cfg, sd, aux, n, noise)
    dt = time.perf_counter() - t0
    return n / dt, n, dt


_REF_MODELS = {}


def cpu_reference_run(kind, kw, frames, sample_frames, threads):
    """Time the UNMODIFIED reference (oracle/_ref copy of src/nets, or /root/reference in the build container):
    ``batch_fast_generate`` on the host CPU, model and inputs on the CPU (the .cuda() identity shim of SURVEY.md 8c),
    wall time around the call as decode_cswnv_laplace-shift1.py:246-249 does, on a bounded sample of the workload."""
    import torch
    from oracle import ref_runner as RR
    torch.set_num_threads(threads)
    sd, aux, n = _cpu_inputs(kind, kw, frames, sample_frames)
    key = (kind, tuple(sorted(kw.items())))
    if key not in _REF_MODELS:
        _REF_MODELS[key] = RR.build_model(kind, kw, sd)
    torch.manual_seed(1)
    out, _, dt = RR.generate(kind, kw, _REF_MODELS[key], aux, [n])
    assert len(out) == 1 and len(out[0]) == (n // (kw.get("seg", 1) if kind == "cswnv" else 1)) * (kw.get("seg", 1) if kind == "cswnv" else 1)
    return n / dt, n, dt


def cpu_arm(kind, kw, frames, sample_frames, threads):
    """(value, n, seconds, kind): the reference itself when its modules are here, else the oracle port (labelled)."""
    from oracle import ref_runner as RR
    if RR.available():
        return cpu_reference_run(kind, kw, frames, sample_frames, threads) + ("reference",)
    return cpu_port_run(kind, kw, frames, sample_frames, threads) + ("port",)


def cpu_baseline_record(kind, kw, frames, sample_frames, receptive_field):
    """cpu_baseline of the B200 line: all host cores and 1 thread, CPU model stated; the oracle
    port is reported beside the reference as a second, labelled figure."""
    import torch
    from oracle import ref_runner as RR
    threads = os.cpu_count() or 1
    v, n, dt, k = cpu_arm(kind, kw, frames, sample_frames, threads)
    v1, n1, dt1, _ = cpu_arm(kind, kw, frames, max(1, sample_frames // 2), 1)
    rec = {"value": v, "unit": "samples/s", "cores": threads, "kind": k,
           "sample": "%s batch_fast_generate on the host CPU, B=1, %d samples (+ the %d-column warm-up), %.1f s wall" % (
               "unmodified reference" if k == "reference" else "oracle port of", n, receptive_field + 1, dt),
           "value_1_thread": v1, "sample_1_thread": "%d samples, %.1f s wall" % (n1, dt1),
           "cpu_model": RR.cpu_model_name(), "nproc": threads, "torch": torch.__version__}
    if k == "reference":
        pv, pn, pdt = cpu_port_run(kind, kw, frames, sample_frames, threads)
        rec["port_value"] = pv
        rec["port_note"] = "oracle/swn_oracle.py (fixed-size history windows instead of the reference's growing torch.cat), same sample"
    torch.set_num_threads(threads)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the batched / sharded sub-records")
    ap.add_argument("--batched-workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--sharded-at-n1", action="store_true", help="also measure the sharded cfg4 sub-record on one GPU (the N=1 figure)")
    ap.add_argument("--sharded-utterances", type=int, default=512)
    ap.add_argument("--cpu-sample-frames", type=int, default=0)
    args = ap.parse_args()

    # stdout carries exactly ONE line (the JSON of rank 0): whatever libraries print there while the run is going on
    # (NCCL's "NCCL version ..." banner when NCCL_DEBUG=VERSION is set on the box) is sent to stderr instead
    sys.stdout.flush()
    _stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kind, fs, over, B, frames, desc = WORKLOADS[args.workload]
    kw = model_kwargs(kind, fs, over)
    U = kw["upsampling_factor"]
    n_samples = frames * U
    macs, n_weights = in_loop_macs(kind, kw)
    metric = "generated samples/sec"
    common = {"metric": metric, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
              "data": "synthetic (random-init weights, synthetic WORLD-style aux, Philox noise)"}
    config = {"workload": desc, "utterances_per_gpu": B, "samples_per_utterance": n_samples, "fs": fs,
              "l2": "flushed between timed steps (256 MiB write)"}

    # ------------------------------------------------------------------ reference arm (CPU port)
    if args.impl == "reference":
        if rank != 0:
            return 0
        import torch
        from oracle import ref_runner as RR
        threads = os.cpu_count() or 1
        # bounded sample per step: the reference generates O(10^2) samples/s at recipe size, so a step is a few
        # frames of the workload (3 frames = 240 samples for DSWNV, 10 frames = 1100 samples for CSWNV seg 5)
        sf = args.cpu_sample_frames or (3 if kind == "dswnv" else 10)
        vals = []
        n = 0
        arm_kind = "port"
        for i in range(args.warmup + args.steps):
            v, n, dt, arm_kind = cpu_arm(kind, kw, frames, sf, threads)
            if i >= args.warmup:
                vals.append((v, dt))
        v1, n1, dt1, _ = cpu_arm(kind, kw, frames, sf, 1)
        v = float(np.mean([x[0] for x in vals]))
        ms = float(np.mean([x[1] for x in vals]) * 1e3)
        sample = ("%s %s.batch_fast_generate on the host CPU, B=1, %d samples (%d frames of the workload) incl. its %d-column "
                  "warm-up, per step" % ("unmodified reference" if arm_kind == "reference" else "oracle port of",
                                         "CSWNV" if kind == "cswnv" else "DSWNV", n, sf, 691))
        line = dict(common)
        line.update({"impl": "reference", "value": v, "ms_per_step": ms, "rtf": v / fs, "config": config,
                     "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": arm_kind, "sample": sample,
                                      "torch_threads": threads, "value_1_thread": v1, "cpu_model": RR.cpu_model_name(),
                                      "nproc": threads, "torch": torch.__version__},
                     "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import shallow_wavenet_b200 as S  # noqa: F401
    from shallow_wavenet_b200 import synth, sharding, engine as E
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a B200; there is no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    smem_peak = E.measure_smem_bandwidth(dev.index or 0)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def reduce_sum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def measure(kind, kw, aux_np, n_samples_list, steps, warmup, e2e_iters, warm_frames=0, last_warm_full=False):
        """Times `steps` passes of the hot path over one batch on this rank: (a) inputs resident in HBM, CUDA events on
        the launching stream, L2 flushed between steps; (b) end to end through batch_fast_generate with pinned host aux
        in and numpy out.  `warm_frames` > 0: warm-up passes run on the first frames only (same batch, same plan)."""
        model = synth.make_model(kind, **kw).to(dev)
        aux_host = torch.from_numpy(aux_np).pin_memory()
        aux_dev = aux_host.to(dev)
        B = aux_np.shape[0]
        seg = kw.get("seg", 1) if kind == "cswnv" else 1
        seed_audio = torch.zeros(B, seg) if kind == "cswnv" else torch.full((B, 1), 128, dtype=torch.long)
        eng = model.engine()
        n_max = max(n_samples_list)
        ragged = len(set(n_samples_list)) > 1
        lens = [int(n) for n in n_samples_list]

        def resident_step(i, frames_limit=0):
            if frames_limit:
                eng.condition(aux_dev[:, :, :frames_limit].contiguous())
                n = frames_limit * kw["upsampling_factor"]
                eng.generate(n, seed=1000 + i, sync=False, n_samples_list=[min(x, n) for x in lens] if ragged else None)
            else:
                eng.condition(aux_dev)
                eng.generate(n_max, seed=1000 + i, sync=False, n_samples_list=lens if ragged else None)

        for i in range(warmup):
            # (short warm-up passes first; the last one at full size, so that no one-off allocation of the full-size
            # conditioning table / outputs falls into the first timed step)
            resident_step(i, 0 if (last_warm_full and i == warmup - 1) else warm_frames)
            eng.sync()
        sampler = ClockSampler(dev.index or 0)
        sampler.start()
        barrier()
        launches0 = eng.get_info()["kernel_launches"]
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        gen_ms, cond_ms = [], []
        for i in range(steps):
            flush.fill_(i & 0xFF)
            ev[i][0].record()
            resident_step(warmup + i)
            ev[i][1].record()
            eng.sync()
            gen_ms.append(eng.info["last_gen_ms"]); cond_ms.append(eng.info["last_cond_ms"])
        barrier()
        launches = eng.get_info()["kernel_launches"] - launches0
        total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
        clocks = sampler.finish()
        e2e_t = []
        if e2e_iters > 0:
            # one untimed call through the public API first: its first use pays one-off host costs (lazy torch / ctypes
            # set-up of this call path, ~50-100 ms measured) that the resident path's warm-up steps never see
            wf_ = warm_frames if warm_frames else aux_host.shape[2]     # (full size where the resident warm-up is full size too)
            model.batch_fast_generate(seed_audio, aux_host[:, :, :wf_], [min(x, wf_ * kw["upsampling_factor"]) for x in lens], seed=1999)
            torch.cuda.synchronize(dev)
        for i in range(e2e_iters):
            barrier()
            t0 = time.perf_counter()
            outs = model.batch_fast_generate(seed_audio, aux_host, lens, seed=2000 + i)
            torch.cuda.synchronize(dev)
            e2e_t.append(time.perf_counter() - t0)
            assert len(outs) == B and all(len(o) == (n // seg) * seg for o, n in zip(outs, lens))
        info = eng.get_info()
        samples = sum((n // seg) * seg for n in lens)
        res = dict(B=B, seg=seg, samples_per_step=samples, n_gen_max=(n_max // seg) * seg, total_ms=total_ms,
                   e2e_ms=float(np.mean(e2e_t)) * 1e3, e2e_ms_each=[t * 1e3 for t in e2e_t], kern_ms=float(np.mean(gen_ms)),
                   cond_ms=float(np.mean(cond_ms)), launches=int(launches), clocks=clocks, info=info,
                   h2d=int(aux_host.numel() * 4), d2h=int(B * (n_max // seg) * seg * (4 if kind == "cswnv" else 8)))
        del eng, model, aux_dev
        torch.cuda.empty_cache()
        return res

    def roofline_of(kind, kw, r, steps):
        """Roofline of the dominant kernel (the persistent generation kernel) of a measured workload."""
        info, kern_ms = r["info"], r["kern_ms"]
        macs_, n_weights_ = in_loop_macs(kind, kw)
        seg_ = r["seg"]
        seg_steps = r["n_gen_max"] // seg_ + info["receptive_field"] // seg_ + 1
        flops = 2.0 * macs_ * r["samples_per_step"] / (kern_ms * 1e-3)
        variant = info.get("kernel_variant")
        kname = {1: "swn_generate_kernel (layer groups over the chip, persistent)",
                 2: "swn_core_kernel (16-CTA core cluster + helper CTAs, persistent)",
                 3: "swn_batch_kernel (ring pipeline of slots; gated cells on tcgen05 kind::tf32 x3, weights in tensor memory)",
                 4: "swn_batch_kernel (ring pipeline of slots; gated cells as register-tiled fp32 FFMA GEMMs)"}.get(variant, "?")
        share = kern_ms * steps / r["total_ms"]
        if variant in (3, 4):
            # batched step: dense contraction [rows x H].[H x columns] -> tensor-pipe bound.  The
            # algorithmic flops count every multiply-add of the reference once (no tf32 pass multiplicity).
            tf = flops / 1e12
            peak_bf16 = peaks.get("bf16_tflops_sustained") or 1400.0
            sm_mhz = r["clocks"].get("sm_mhz") or 1965.0
            fp32_peak = info["n_sm"] * 128 * 2 * sm_mhz * 1e6 / 1e12
            return {"bound": "tensor", "kernel": kname + ", 1 launch per step",
                    "achieved": tf, "peak": peak_bf16, "unit": "TFLOP/s", "frac": tf / peak_bf16,
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16, long loop)" if peaks.get("bf16_tflops_sustained")
                                    else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"),
                    "note": "fp32 parity (1e-4) forbids bf16: the cells run three tf32 passes (tf32 dense = bf16/2), so the "
                            "same-precision tensor ceiling is peak/6; variant 4 (H=256 does not fit tensor memory) is bounded by fp32 FFMA",
                    "peak_tf32x3_equivalent": peak_bf16 / 6.0, "frac_of_tf32x3_equivalent": tf / (peak_bf16 / 6.0),
                    "fp32_ffma_peak_at_clock": fp32_peak, "frac_of_fp32_ffma_peak": tf / fp32_peak,
                    "algorithmic_flop_per_sample": 2.0 * macs_, "samples_per_launch": r["samples_per_step"],
                    "kernel_ms": kern_ms, "kernel_share_of_step": share,
                    "us_per_loop_iteration": kern_ms * 1e3 / seg_steps, "traffic": None}
        weight_bytes = n_weights_ * 4.0
        alg_bytes = weight_bytes * seg_steps * info["n_slots"]
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        return {"bound": "smem", "kernel": kname + ", 1 launch per step",
                "achieved": achieved, "peak": smem_peak, "unit": "GB/s", "frac": achieved / smem_peak,
                "peak_source": "measured live: aggregate LDS.128 bandwidth, 1 CTA/SM (swn_measure_smem_bandwidth)",
                "algorithmic_bytes_per_time_step": weight_bytes, "time_steps_per_launch": seg_steps * info["n_slots"],
                "us_per_time_step": kern_ms * 1e3 / (seg_steps * info["n_slots"]),
                "kernel_ms": kern_ms, "kernel_share_of_step": share,
                "achieved_tflops_fp32": flops / 1e12,
                "vs_hbm": {"peak": peaks.get("hbm_gbs"), "frac": (achieved / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                           "note": "weights are SMEM-resident; HBM is not the bound (frac>1 = residency pays)"},
                "traffic": None}

    def attach_traffic(roofline, workload, variant):
        # DRAM bytes per launch of the dominant kernel from one `ncu --set full` capture of this workload
        # (dram__bytes_read.sum + dram__bytes_write.sum), recorded under profiles/; null when none was taken
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(workload)
            if cap and cap.get("kernel_variant") == variant:
                roofline["traffic"] = cap["dram_bytes_per_launch"]
                roofline["traffic_source"] = cap["source"]
        except Exception:
            pass

    # ---- the headline workload: every rank generates its own copy (weak scaling, no collective on the data path)
    aux_np = synth.synth_aux(kw["n_aux"], frames, B, seed=1234 + rank)
    r = measure(kind, kw, aux_np, [n_samples] * B, args.steps, args.warmup, max(1, min(args.steps, 3)))
    total_ms, e2e_ms = reduce_max([r["total_ms"], r["e2e_ms"]])
    samples_all = world * r["samples_per_step"]
    value = samples_all * args.steps / (total_ms * 1e-3)
    e2e_value = samples_all / (e2e_ms * 1e-3)
    line = None
    if rank == 0:
        roofline = roofline_of(kind, kw, r, args.steps)
        attach_traffic(roofline, args.workload, r["info"].get("kernel_variant"))
        line = dict(common)
        line.update({"value": value, "ms_per_step": total_ms / args.steps, "rtf": value / fs, "config": config,
                     "rtf_per_utterance": value / fs / (world * B),
                     "conditioning_ms": r["cond_ms"], "kernel_ms": r["kern_ms"],
                     "clocks": r["clocks"], "roofline": roofline,
                     "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                             "ms_each": r["e2e_ms_each"],
                             "api": "%s.batch_fast_generate(pinned host aux) -> list of numpy arrays" % ("CSWNV" if kind == "cswnv" else "DSWNV")},
                     "gpu_launches": r["launches"],
                     "plan": {k: r["info"][k] for k in ("kernel_variant", "grid", "n_sm", "ctas_per_layer", "head_ctas", "smem_bytes",
                                                        "batch_tile", "n_slots")}})

    # ---- batched sub-record: BASELINE.json configs[2] (tensor-core batched step), measured in the same run
    if not args.no_extras and args.workload == "cfg2":
        bk, bfs, bover, bB, bframes, bdesc = WORKLOADS[args.batched_workload]
        bkw = model_kwargs(bk, bfs, bover)
        bn = bframes * bkw["upsampling_factor"]
        rb = measure(bk, bkw, synth.synth_aux(bkw["n_aux"], bframes, bB, seed=4321 + rank), [bn] * bB, 2, 3, 1, warm_frames=40, last_warm_full=True)
        b_total, b_e2e = reduce_max([rb["total_ms"], rb["e2e_ms"]])
        if rank == 0:
            bro = roofline_of(bk, bkw, rb, 2)
            attach_traffic(bro, args.batched_workload, rb["info"].get("kernel_variant"))
            bval = world * rb["samples_per_step"] * 2 / (b_total * 1e-3)
            line["batched"] = {args.batched_workload: {
                "workload": bdesc, "value": bval, "unit": "samples/s", "rtf": bval / bfs, "rtf_per_utterance": bval / bfs / (world * bB),
                "steps": 2, "warmup": "2 passes on the first 40 frames (same batch, same plan) + 1 at full size", "scaling": "weak",
                "ms_per_step": b_total / 2, "kernel_ms": rb["kern_ms"], "conditioning_ms": rb["cond_ms"],
                "e2e": {"value": world * rb["samples_per_step"] / (b_e2e * 1e-3), "unit": "samples/s",
                        "h2d_bytes_per_step": rb["h2d"], "d2h_bytes_per_step": rb["d2h"]},
                "roofline": bro, "clocks": rb["clocks"], "gpu_launches": rb["launches"],
                "plan": {k: rb["info"][k] for k in ("kernel_variant", "grid", "batch_tile", "n_slots", "ctas_per_layer", "head_ctas")}}}

    # ---- latency sub-record: BASELINE.json configs[0] on the GPU (the Laplace recipe model at batch 1; the seg > 1 cluster
    # kernel, swn_coreseg.cuh), measured in the same run
    if not args.no_extras and args.workload == "cfg2":
        lk, lfs, lover, lB, lframes, ldesc = WORKLOADS["cfg1"]
        lkw = model_kwargs(lk, lfs, lover)
        ln = lframes * lkw["upsampling_factor"]
        rl = measure(lk, lkw, synth.synth_aux(lkw["n_aux"], lframes, lB, seed=2345 + rank), [ln] * lB, 3, 2, 3)
        l_total, l_e2e = reduce_max([rl["total_ms"], rl["e2e_ms"]])
        if rank == 0:
            lro = roofline_of(lk, lkw, rl, 3)
            attach_traffic(lro, "cfg1", rl["info"].get("kernel_variant"))
            lval = world * rl["samples_per_step"] * 3 / (l_total * 1e-3)
            line["latency"] = {"cfg1": {
                "workload": ldesc, "value": lval, "unit": "samples/s", "rtf": lval / lfs, "rtf_per_utterance": lval / lfs / (world * lB),
                "steps": 3, "warmup": 2, "scaling": "weak", "ms_per_step": l_total / 3, "kernel_ms": rl["kern_ms"],
                "conditioning_ms": rl["cond_ms"],
                "e2e": {"value": world * rl["samples_per_step"] / (l_e2e * 1e-3), "unit": "samples/s",
                        "h2d_bytes_per_step": rl["h2d"], "d2h_bytes_per_step": rl["d2h"]},
                "roofline": lro, "clocks": rl["clocks"], "gpu_launches": rl["launches"],
                "plan": {k: rl["info"][k] for k in ("kernel_variant", "grid", "batch_tile", "n_slots", "ctas_per_layer", "head_ctas")}}}

    # ---- decode-batch sub-record: the cfg1 model at the recipe's default decode batch of 7 utterances (run.sh:280): seg cluster
    # kernel, three utterances side by side per launch (3 + 2 + 2)
    for dname in (("cfg1x7", "seg1x7", "cfg2x7") if (not args.no_extras and args.workload == "cfg2") else ()):
        dk, dfs, dover, dB, dframes, ddesc = WORKLOADS[dname]
        dkw = model_kwargs(dk, dfs, dover)
        dn = dframes * dkw["upsampling_factor"]
        rd = measure(dk, dkw, synth.synth_aux(dkw["n_aux"], dframes, dB, seed=3456 + rank), [dn] * dB, 3, 2, 3)
        d_total, d_e2e = reduce_max([rd["total_ms"], rd["e2e_ms"]])
        if rank == 0:
            dro = roofline_of(dk, dkw, rd, 3)
            dval = world * rd["samples_per_step"] * 3 / (d_total * 1e-3)
            line.setdefault("decode_batch", {})[dname] = {
                "workload": ddesc, "value": dval, "unit": "samples/s", "rtf": dval / dfs, "rtf_per_utterance": dval / dfs / (world * dB),
                "steps": 3, "warmup": 2, "scaling": "weak", "ms_per_step": d_total / 3, "kernel_ms": rd["kern_ms"],
                "conditioning_ms": rd["cond_ms"],
                "e2e": {"value": world * rd["samples_per_step"] / (d_e2e * 1e-3), "unit": "samples/s",
                        "h2d_bytes_per_step": rd["h2d"], "d2h_bytes_per_step": rd["d2h"]},
                "roofline": dro, "clocks": rd["clocks"], "gpu_launches": rd["launches"],
                "plan": {k: rd["info"][k] for k in ("kernel_variant", "grid", "batch_tile", "n_slots", "ctas_per_layer", "head_ctas")}}

    # ---- sharded sub-record: BASELINE.json configs[3], 512 utterances of unequal length split over the ranks the way
    # the decode scripts split their list (contiguous np.array_split, decode_cswnv_laplace-shift1.py:200-201), each
    # rank sorting its share by length (:77-79) and decoding it as one batch.  STRONG scaling: total work is fixed.
    if not args.no_extras and args.workload == "cfg2" and (world > 1 or args.sharded_at_n1):
        sk, sfs, sover, _, sframes, sdesc = WORKLOADS["cfg4"]
        skw = model_kwargs(sk, sfs, sover)
        U_s = skw["upsampling_factor"]
        n_utt = args.sharded_utterances
        lens_all = np.random.default_rng(77).integers(int(0.8 * sframes), int(1.2 * sframes) + 1, n_utt)    # frames, +-20 %
        mine = [int(i) for i in sharding.rank_shard(list(range(n_utt)), rank, world)]
        order = [mine[i] for i in sharding.decode_batches([lens_all[i] for i in mine], len(mine))[0]]
        T_max = int(max(lens_all[i] for i in order))
        aux_s = np.zeros((len(order), skw["n_aux"], T_max), np.float32)
        for row, i in enumerate(order):
            aux_s[row, :, :lens_all[i]] = synth.synth_aux(skw["n_aux"], int(lens_all[i]), 1, seed=9000 + i)[0]
        n_list = [int(lens_all[i]) * U_s for i in order]
        rs = measure(sk, skw, aux_s, n_list, 1, 1, 1, warm_frames=40)
        s_total, s_e2e = reduce_max([rs["total_ms"], rs["e2e_ms"]])
        s_sum_ms, s_samples = reduce_sum([rs["total_ms"], float(rs["samples_per_step"])])
        if rank == 0:
            sval = s_samples / (s_total * 1e-3)
            sro = roofline_of(sk, skw, rs, 1)
            rec = {"workload": "cfg4: CSWNV Laplace seg=5 lpc=4 H=192, 22.05 kHz, %d utterances of 8-12 s (10 s +-20 %%) sharded over "
                               "%d GPU(s) with sharding.split_for_gpus; each rank decodes its share as one length-sorted batch" % (n_utt, world),
                   "value": sval, "unit": "samples/s", "rtf": sval / sfs, "scaling": "strong", "steps": 1,
                   "warmup": "1 pass on the first 40 frames (same batch, same plan)",
                   "utterances_total": n_utt, "utterances_rank0": len(order), "samples_total": s_samples,
                   "ms_max_over_ranks": s_total, "ms_mean_over_ranks": s_sum_ms / world, "imbalance_max_over_mean": s_total * world / s_sum_ms,
                   "e2e": {"value": s_samples / (s_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step_rank0": rs["h2d"],
                           "d2h_bytes_per_step_rank0": rs["d2h"]},
                   "rank0": {"kernel_ms": rs["kern_ms"], "conditioning_ms": rs["cond_ms"], "roofline": sro,
                             "conditioning_table_bytes": rs["info"]["cond_bytes"],
                             "plan": {k: rs["info"][k] for k in ("kernel_variant", "grid", "batch_tile", "n_slots")}}}
            try:
                n1 = json.load(open(os.path.join(ROOT, "profiles", "r02_sharded_cfg4_n1.json")))
                if n1.get("utterances_total") == n_utt:
                    rec["n1_value"] = n1["value"]
                    rec["n1_source"] = "profiles/r02_sharded_cfg4_n1.json (same 512 utterances on one GPU)"
                    rec["efficiency_vs_n1"] = sval / (world * n1["value"])
            except Exception:
                pass
            line["sharded"] = {"cfg4": rec}

    if rank == 0:
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only (the other ranks would spin in the barrier)
            sf = args.cpu_sample_frames or (8 if kind == "dswnv" else 24)
            line["cpu_baseline"] = cpu_baseline_record(kind, kw, frames, sf, r["info"]["receptive_field"])
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
