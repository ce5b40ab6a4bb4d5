#!/usr/bin/env python
"""Benchmark of the keyword-spotting hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, BASELINE config 3 (default)
  python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference path
  python bench.py --config 2|3|4|5 [--job]                 # the other BASELINE configs / the fixed 158,538-clip job

Default workload (config.workload): BASELINE.json configs[2] = exp-195 Depthwise1D forward with 8x TTA, with
the north-star stages in front of it -- one "step" is one pass over a batch of B synthetic 1 s / 16 kHz
clips: augment (time-shift + noise mix + volumes) -> log-mel (40 mel, 30 ms / 10 ms) -> 8-view forward
-> TTA mean -> argmax.  B clips x 64 KB = 1.05 GB at the default B=16384 (about the per-GPU share of
the 158,538-clip job on 8 GPUs), far larger than the 126 MB L2; the forward runs in chunks of
--max-rows clip-views (65536 = 8192 clips: two chunks per step; r01 used 32768).
Per-GPU work is fixed as N grows (weak scaling); the only collective is one all-gather of the
[B,12] probabilities per step.  `value` times the device-resident path with CUDA events on the
launching stream (max over ranks); `e2e` times the host-buffer C-ABI call with the same work: clips in the
reference's wire format (int16 PCM, input_data.py:334-336) in pinned host memory, H2D + D2H inside, probabilities
and labels read back (`e2e_variants` holds the fp32-input / feature-readback / pageable-caller forms).

Other configs (one JSON line each, same contract; kept under profiles/):
  --config 2        log-mel front end only (augment + STFT + mel + log), B = 4096 on one GPU
  --config 3 --job  the fixed job: all 158,538 clips as ONE step, sharded over the ranks (strong scaling)
  --config 4        exp-106 net (32 classes) x 3 TTA views -> 32->12 map + re-softmax + uint8 -> threshold 0.6
  --config 5        106 + 195 + 206 x 3 TTA views each -> 3-way majority vote, batch sweep 1k .. 64k
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CLIPS_JOB = 158538                       # convert_from_see_v3_bugfix.py:66
FLOP_PER_VIEW = {195: 112.48e6, 106: 98.04e6}
FLOP_CONV1, FLOP_HEAD = 12.26e6, 0.11e6
FLOP_BLOCKS = FLOP_PER_VIEW[195] - FLOP_CONV1 - FLOP_HEAD
FLOP_FRONTEND = 50.7e6                     # DFT-as-GEMM + mel + DCT per clip
BYTES_AUGMENT = 192020                     # per clip
BYTES_FRONTEND = 143700                    # per clip, config 2 (wav + noise slice + params + log-mel out)
ACT_BYTES_VIEW_FP16 = 286720 * 2 * 2       # 12 block outputs written + read once, fp16
METRIC = "1s-clips/sec (aug+feat+fwd, 8x TTA)"
WORKLOADS = {
    2: "config2: augment + log-mel front end only (40 mel, 30/10 ms window/hop); one step = one batch of 4096 clips",
    3: ("config3: augment + log-mel(40 mel, 30/10 ms) + exp-195 Depthwise1D forward x 8 TTA views + TTA mean "
        "+ argmax; one step = one batch"),
    4: ("config4: exp-106 Depthwise1D forward x 3 TTA views + TTA mean -> 32->12 max-map + re-softmax + uint8 "
        "-> pseudo-label threshold 0.6; one step = one batch"),
    5: ("config5: exp-106 + 195 + 206 forwards x 3 TTA views each + TTA means + argmax -> 3-way majority vote "
        "(min_count 2, model-0 fallback); one step = one batch; batch sweep 1k..64k"),
}


def block_bytes_per_view(arch=195):
    """Algorithmic HBM bytes of each depthwise+pointwise block kernel per clip-view in the
    block-materialised model of SURVEY.md 8(d) with fp16 activations: the block reads its input
    activation [t_in, cin] once and writes its output [t_out, cout] once."""
    from speech_recognition_b200.arch import ARCHS, layer_lengths
    a = ARCHS[arch]
    T = layer_lengths(arch)[1:]
    out, cin = [], a["conv1"]
    for i, (co, _) in enumerate(a["blocks"]):
        out.append((T[i] * cin + T[i + 1] * co) * 2)
        cin = co
    return out


def measured_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    tflops_burst=p["bf16_tflops"], source="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1590.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, reasons, smax = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return None
        busy = [c for c in sm if c > 0]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def pin_to_gpu_numa(local):
    """Bind this rank (and the pinned host memory it allocates afterwards: first touch) to the CPUs of its GPU's
    NUMA node, so that 8 ranks do not all pull their H2D traffic out of node 0.  Returns a description."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        base = f"/sys/bus/pci/devices/{bus}"
        with open(base + "/local_cpulist") as f:
            txt = f.read().strip()
        node = open(base + "/numa_node").read().strip() if os.path.exists(base + "/numa_node") else "?"
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"pci": bus, "numa_node": node, "cpus": txt, "bound": bool(allowed)}
    except Exception as e:                                       # noqa: BLE001
        return {"bound": False, "why": str(e)[:80]}


# --------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle), timed on the host cores
# --------------------------------------------------------------------------------------------
class CpuReference:
    """augment + log-mel + n-view forward of one batch with the oracle on all host threads.
    bench.py may execute oracle/ only here (cpu_baseline / --impl reference)."""

    def __init__(self, views, batch=64, config=3):
        import torch
        from oracle import augment, frontend, network, driver
        from speech_recognition_b200 import synth
        from speech_recognition_b200.classes import class_map_32_to_12
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.batch, self.views, self.config = batch, views, config
        clips = synth.make_clips(batch, seed=synth.SEED + 3)
        bank, offs = synth.make_noise_bank(seconds=4)
        p = synth.make_params(batch, offs, seed=synth.SEED + 4)
        archs = {2: (), 3: (195,), 4: (106,), 5: (106, 195, 206)}[config]
        ws = {a: synth.synthetic_weights(a) for a in archs}

        def one_batch():
            bg = augment.gather_background(bank, offs, p["bg_index"], p["bg_offset"])
            x = augment.augment_mix(clips, p["time_shift"], bg, p["bg_volume"], p["fg_volume"])
            if config in (2, 3):
                frontend.features(x, kind="logmel", dct_coefficient_count=40, fft_dtype=np.float32)
            labels = []
            for a in archs:
                pr, lab = driver.tta_predict(lambda v: network.forward(v, ws[a], a, dtype=torch.float32), x, views)
                labels.append(lab)
                if config == 4:
                    _, u8 = driver.convert_32_to_12(pr.astype(np.float32), "heng")
                    driver.threshold_select(u8, 0.6)
            if config == 5:
                cm = np.asarray(class_map_32_to_12("frozen"))
                driver.majority_vote(np.stack([cm[l] if a == 106 else l for a, l in zip(archs, labels)]).astype(np.int32), 2)
        self.one_batch = one_batch

    def rate(self, seconds_target=15.0, max_batches=16):
        self.one_batch()                                         # warm-up (oneDNN primitive caches)
        t0 = time.perf_counter(); self.one_batch(); t1 = time.perf_counter() - t0
        n = int(max(1, min(max_batches, seconds_target / max(t1, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(n):
            self.one_batch()
        dt = time.perf_counter() - t0
        return dict(value=self.batch * n / dt, unit="clips/s", cores=self.threads, kind="port",
                    sample=f"{n} batches of {self.batch} clips x {len(self.views)} views of config {self.config}, "
                           "torch-CPU fp32 oracle")


def views_of(config):
    from speech_recognition_b200 import TTA_8, TTA_SHIPPED
    return TTA_8 if config == 3 else (TTA_SHIPPED if config in (4, 5) else ())


def run_reference(args):
    """The reference arm: the CPU restatement on a BOUNDED sample of the workload.  One reference step = one batch
    of 64 clips (BASELINE config 1's batch) through the same stages; W warm-up and exactly K timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    views = views_of(args.config)
    ref = CpuReference(views, batch=64, config=args.config)
    t_wall = time.perf_counter()
    for _ in range(max(1, args.warmup)):
        ref.one_batch()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.one_batch()
    dt = time.perf_counter() - t0
    value = 64 * args.steps / dt
    cpu = dict(value=value, unit="clips/s", cores=ref.threads, kind="port",
               sample=f"{args.steps} steps of 64 clips x {len(views)} views, torch-CPU fp32 oracle, {ref.threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(1, args.warmup),
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "batch_per_gpu": 64, "views": len(views), "job_clips": N_CLIPS_JOB,
                   "reference_arm": "CPU restatement of the reference path (oracle/, torch-CPU fp32, all host threads; "
                                    "TF 1.4 / Keras 2.1.2 are not installable here).  Bounded sample of the workload: one "
                                    "step = one batch of 64 clips (the repo arm's step is 16384 clips per GPU); clips/s "
                                    "is the comparable quantity, ms_per_step is not"},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_wall,
    }
    emit(line)


# --------------------------------------------------------------------------------------------
class Ctx:
    """Per-rank setup shared by the configs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.numa = pin_to_gpu_numa(self.local) if not args.no_numa else {"bound": False, "why": "--no-numa"}
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local}"))
        self.dev = torch.device(f"cuda:{self.local}")
        self.args = args
        # clips per staged chunk of the *_host entry points (libkws reads the variable once): 8,192-clip chunks are ~3 %
        # faster on one or two GPUs (larger launches), 4,096-clip ones from four GPUs on, where the ranks share the
        # host's memory bandwidth and the exposed first upload / last download grows with the chunk (r02n / r02r A/B)
        # (config 3 only: the copy-bound 3-view configs 4 / 5 lose 10-15 % to the larger chunk)
        os.environ.setdefault("KWS_HOST_CHUNK", "8192" if (self.world <= 2 and args.config == 3) else "4096")
        self.host_chunk = int(os.environ["KWS_HOST_CHUNK"])

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, steps, warmup):
        """W warm-up steps, barrier + sync, exactly K timed steps between CUDA events, barrier + sync; max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        self.sync_all()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def timed_host(self, step, steps):
        """Host-clocked region for the host-buffer entry points (they return when the results are in host memory)."""
        step()
        self.sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.torch.cuda.synchronize()
        return self.max_over_ranks(time.perf_counter() - t0)

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def make_inputs(ctx, B, offs, pool_size=512):
    """A pool of distinct synthetic clips tiled to the batch (every rank a different slice) + pre-drawn parameters."""
    from speech_recognition_b200 import synth
    torch = ctx.torch
    clips, pcm = synth.make_clips(pool_size, seed=synth.SEED + 17 * ctx.rank, return_pcm=True)
    reps = (B + pool_size - 1) // pool_size
    x = torch.from_numpy(clips).to(ctx.dev).repeat(reps, 1)[:B].contiguous()
    xi = torch.from_numpy(pcm).repeat(reps, 1)[:B].contiguous()               # host int16
    p = synth.make_params(B, offs, seed=synth.SEED + 1 + ctx.rank)
    return x, xi, p


def pinned(ctx, t):
    out = ctx.torch.empty(t.shape, dtype=t.dtype).pin_memory()
    out.copy_(t)
    return out


def base_line(ctx, args, value, ms, steps, config_extra, dtype="f16"):
    return {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": ctx.world, "steps": steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype if args.precision == "tc" else "f32", "data": "synthetic",
        "config": dict({"workload": WORKLOADS[args.config], "precision": args.precision, "max_rows": args.max_rows,
                        "job_clips": N_CLIPS_JOB, "numa": ctx.numa, "host_chunk_clips": ctx.host_chunk}, **config_extra),
    }


# -------------------------------------------------------------------------------------------- config 3
def run_config3(args):
    from speech_recognition_b200 import Engine, synth, TTA_8
    from speech_recognition_b200.sharded import shard_range, all_gather_rows
    ctx = Ctx(args)
    torch, dist, dev, world, rank = ctx.torch, ctx.dist, ctx.dev, ctx.world, ctx.rank
    V, views = 8, TTA_8
    if args.job:
        s, e = shard_range(N_CLIPS_JOB, world, rank)
        B = e - s
    else:
        B = args.batch
    eng = Engine(device=ctx.local, max_rows=args.max_rows, precision=args.precision)
    bank, offs = synth.make_noise_bank(seconds=60)                 # 23 MB, resident per GPU
    eng.set_noise_bank(torch.from_numpy(bank).to(dev), offs)
    eng.frontend_config(480, 160, 40, 40)                          # BASELINE config 2 front end
    eng.load_model(0, 195, synth.synthetic_weights(195))
    x, xi, p = make_inputs(ctx, B, offs)
    pt = {k: torch.from_numpy(v).to(dev) for k, v in p.items()}
    aug = torch.empty_like(x)
    feat = torch.empty((B, 98, 40), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * B, 12), dtype=torch.float32, device=dev) if (world > 1 and not args.job) else None

    def step():
        eng.augment(x, pt["time_shift"], pt["bg_index"], pt["bg_offset"], pt["bg_volume"], pt["fg_volume"], out_t=aug)
        eng.features(aug, "logmel", out_t=feat)
        probs, amax = eng.forward(aug, views=views)
        if args.job:
            return all_gather_rows(probs, N_CLIPS_JOB), amax      # one padded all-gather (uneven last shard)
        if world > 1:
            dist.all_gather_into_tensor(gathered, probs)
        return probs, amax

    # ---- pass A: the headline, no per-launch event brackets ----
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    ctx.sync_all()
    if sampler:
        sampler.start()
    launches0 = eng.launch_count
    ms = ctx.timed(step, args.steps, 0)
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count - launches0
    total_clips = N_CLIPS_JOB if args.job else world * B
    value = total_clips * args.steps / (ms / 1e3)
    # ---- pass B: per-kernel-class device timing for the roofline (events around every launch) ----
    eng.timing_read(); eng.timing_enable(True)
    t_steps = max(2, min(args.steps, 5))
    for _ in range(t_steps):
        step()
    classes = eng.timing_read()
    eng.timing_enable(False)

    # ---- e2e through the host-buffer C-ABI entry point ----
    e2e_steps = max(1, min(args.steps, 5))
    hp = {k: pinned(ctx, torch.from_numpy(v)).numpy() for k, v in p.items()}
    h_probs = torch.empty((B, 12), dtype=torch.float32).pin_memory()
    h_amax = torch.empty((B,), dtype=torch.int32).pin_memory()
    h_feat = torch.empty((B, 98 * 40), dtype=torch.float32).pin_memory()
    hx16 = pinned(ctx, xi)
    par_bytes = 5 * B * 4
    out_bytes = B * 12 * 4 + B * 4

    def gather_host():
        if args.job:
            all_gather_rows(h_probs.to(dev, non_blocking=True), N_CLIPS_JOB)

    def e2e_pcm():
        eng.pipeline_host(hx16.numpy(), hp, feat_kind="logmel", views=views, want_features=False,
                          probs_out=h_probs.numpy(), argmax_out=h_amax.numpy())
        gather_host()
    dt = ctx.timed_host(e2e_pcm, e2e_steps)
    e2e_value = total_clips * e2e_steps / dt
    variants = {}
    if not args.job and not args.quick:
        hx32 = pinned(ctx, x.cpu())

        def e2e_f32_feat():
            eng.pipeline_host(hx32.numpy(), hp, feat_kind="logmel", views=views, feat_out=h_feat.numpy(),
                              probs_out=h_probs.numpy(), argmax_out=h_amax.numpy())

        def e2e_pcm_feat():
            eng.pipeline_host(hx16.numpy(), hp, feat_kind="logmel", views=views, feat_out=h_feat.numpy(),
                              probs_out=h_probs.numpy(), argmax_out=h_amax.numpy())
        pg16, pg_par = xi.numpy().copy(), {k: v.copy() for k, v in p.items()}      # plain (pageable) np.ndarrays

        def e2e_pageable():
            eng.pipeline_host(pg16, pg_par, feat_kind="logmel", views=views, want_features=False)
        for name, fn, h2d, d2h in (("fp32_pinned_in_features_out (r01 e2e)", e2e_f32_feat, B * 64000 + par_bytes, out_bytes + B * 15680),
                                   ("pcm16_pinned_in_features_out", e2e_pcm_feat, B * 32000 + par_bytes, out_bytes + B * 15680),
                                   ("pcm16_pageable_ndarray_in (staged through pinned slots)", e2e_pageable, B * 32000 + par_bytes, out_bytes)):
            d = ctx.timed_host(fn, e2e_steps)
            variants[name] = {"value": world * B * e2e_steps / d, "unit": "clips/s", "h2d_bytes_per_step": h2d,
                              "d2h_bytes_per_step": d2h}
        del hx32

    if rank == 0:
        peaks = load_peaks()
        blk_ms, blk_n = classes["dw_pw_blocks"]
        views_per_step = B * V
        bpv = block_bytes_per_view(195)
        views_per_launch = min(args.max_rows // V * V, views_per_step)
        per_block = getattr(eng, "last_block_ms", None)
        active = [i for i in range(len(bpv)) if per_block and per_block[i][1] > 0] or list(range(len(bpv)))
        # launches of a step differ in size when B is not a multiple of the chunk: bytes / time over the whole pass
        alg_bytes_pass = sum(bpv[i] for i in active) * views_per_step * t_steps
        blk_gbs = alg_bytes_pass / (blk_ms / 1e3) / 1e9 if blk_ms > 0 else 0.0
        blk_avg_launch_ms = blk_ms / max(blk_n, 1)
        alg_bytes_per_launch = alg_bytes_pass / max(blk_n, 1)
        blk_flop = FLOP_BLOCKS * sum(bpv[i] for i in active) / sum(bpv)
        blk_tflops = blk_flop * views_per_step * t_steps / (blk_ms / 1e3) / 1e12 if blk_ms > 0 else 0.0
        fused_first = bool(per_block) and per_block[0][1] == 0
        tr = measured_traffic()
        t_total = sum(v[0] for v in classes.values())
        per_class = {k: {"ms_per_step": v[0] / t_steps, "launches_per_step": v[1] / t_steps}
                     for k, v in classes.items() if v[1]}
        aug_ms = classes["augment"][0] / t_steps
        line = base_line(ctx, args, value, ms, args.steps, {
            "batch_per_gpu": B, "views": V,
            "l2_policy": "inputs larger than L2 (batch x 64 KB = %.0f MB)" % (B * 64e3 / 1e6),
            "fused_conv1_block1": fused_first,
            "parallelism": (f"dp{world} (the 158,538-clip job sharded 19,818 x 7 + 19,812; one padded all-gather of probabilities)"
                            if args.job else f"dp{world} (clip shards, 1 all-gather of probabilities per step)")})
        if args.job:
            line["scaling"] = "strong"
            line["config"]["workload"] += "; --job: all 158,538 clips (convert_from_see_v3_bugfix.py:66) as ONE step"
        line.update({
            "roofline": {"kernel": ("tc_gemm_kernel<1|2> (TMA-fed depthwise producer + tcgen05 pointwise GEMM + BN/ReLU6 "
                                    "+ TMA store), %d launches per chunk" % len(active)) if args.precision == "tc" else "gemm_f32_kernel",
                         "bound": "hbm", "achieved": blk_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": blk_gbs / peaks["hbm_gbs"], "peak_source": peaks["source"] + " (copy bandwidth)",
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                         "avg_launch_ms": blk_avg_launch_ms, "clip_views_per_launch": views_per_launch,
                         # the ncu capture was taken at 32,768 clip-views per launch; traffic is linear in the rows
                         "traffic": tr["bytes_per_launch"] * views_per_launch / tr.get("clip_views_per_launch", views_per_launch) if tr else None,
                         "traffic_source": (tr["source"] + "; scaled from %d to %d clip-views per launch" % (
                             tr.get("clip_views_per_launch", views_per_launch), views_per_launch)) if tr else None,
                         "share_of_step": blk_ms / t_total if t_total else None,
                         "tensor_tflops": blk_tflops, "tensor_frac": blk_tflops / peaks["tflops"]},
            "secondary_rooflines": {
                "augment_hbm_gbs": BYTES_AUGMENT * B / (aug_ms / 1e3) / 1e9 if aug_ms else None,
                "augment_hbm_frac": (BYTES_AUGMENT * B / (aug_ms / 1e3) / 1e9) / peaks["hbm_gbs"] if aug_ms else None,
                "whole_step_tensor_frac": (FLOP_PER_VIEW[195] * V + FLOP_FRONTEND) * B * args.steps / (ms / 1e3) / 1e12 / peaks["tflops"],
                "whole_step_hbm_frac_block_materialised": (ACT_BYTES_VIEW_FP16 * V + BYTES_AUGMENT + 64000 + 15680) * B
                * args.steps / (ms / 1e3) / 1e9 / peaks["hbm_gbs"],
            },
            "kernel_classes": per_class,
            "kernel_timing_pass": {"steps": t_steps, "note": "separate pass with CUDA events around every launch; the "
                                   "headline pass above runs without them"},
            "block_ms_per_step": [round(b[0] / t_steps, 4) for b in per_block] if per_block else None,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": B * 32000 + par_bytes,
                    "d2h_bytes_per_step": out_bytes, "steps": e2e_steps,
                    "api": "kws_pipeline_host_pcm16 (ctypes): int16 PCM clips + parameters in pinned host memory in, "
                           "probabilities + labels out; log-mel computed on the device and left there (the shipped "
                           "networks consume the raw waveform, make_submission.py:46)"},
            "e2e_variants": variants,
            "gpu_launches": int(launches),
            "clocks": clocks,
        })
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = CpuReference(views, config=3).rate(seconds_target=15.0)
        emit(line)
    ctx.finish()
    eng.close()


# -------------------------------------------------------------------------------------------- config 2
def run_config2(args):
    from speech_recognition_b200 import Engine, synth
    ctx = Ctx(args)
    torch, dev, world, rank = ctx.torch, ctx.dev, ctx.world, ctx.rank
    B = args.batch if args.batch_given else 4096
    eng = Engine(device=ctx.local, max_rows=args.max_rows, precision=args.precision)
    bank, offs = synth.make_noise_bank(seconds=60)
    eng.set_noise_bank(torch.from_numpy(bank).to(dev), offs)
    eng.frontend_config(480, 160, 40, 40)
    x, xi, p = make_inputs(ctx, B, offs, pool_size=min(B, 4096))           # 4096 distinct clips: 262 MB > L2
    pt = {k: torch.from_numpy(v).to(dev) for k, v in p.items()}
    aug = torch.empty_like(x)
    feat = torch.empty((B, 98, 40), dtype=torch.float32, device=dev)

    def step():
        eng.augment(x, pt["time_shift"], pt["bg_index"], pt["bg_offset"], pt["bg_volume"], pt["fg_volume"], out_t=aug)
        eng.features(aug, "logmel", out_t=feat)
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    for _ in range(3):
        step()
    ctx.sync_all()
    if sampler:
        sampler.start()
    launches0 = eng.launch_count
    ms = ctx.timed(step, args.steps, 0)
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    value = world * B * args.steps / (ms / 1e3)
    eng.timing_read(); eng.timing_enable(True)
    for _ in range(5):
        step()
    classes = eng.timing_read(); eng.timing_enable(False)
    hp = {k: pinned(ctx, torch.from_numpy(v)).numpy() for k, v in p.items()}
    hx16 = pinned(ctx, xi)
    h_feat = torch.empty((B, 98 * 40), dtype=torch.float32).pin_memory()

    def e2e():
        eng.pipeline_host(hx16.numpy(), hp, feat_kind="logmel", views=(), feat_out=h_feat.numpy())
    e2e_steps = max(1, min(args.steps, 10))
    dt = ctx.timed_host(e2e, e2e_steps)
    if rank == 0:
        peaks = load_peaks()
        dft_ms, dft_n = classes["dft"]
        aug_ms = classes["augment"][0] / 5
        fe_tflops = FLOP_FRONTEND * B * 5 / (dft_ms / 1e3) / 1e12 if dft_ms else 0.0
        line = base_line(ctx, args, value, ms, args.steps, {
            "batch_per_gpu": B, "l2_policy": "inputs larger than L2 (%d distinct clips x 64 KB = %.0f MB)" % (min(B, 4096), min(B, 4096) * 64e3 / 1e6),
            "parallelism": f"dp{world} (replicas, no collective)"}, dtype="f16x3 (split-fp16, fp32-accurate)")
        line.update({
            "roofline": {"kernel": "stft_mel_tc_kernel (tcgen05 DFT-as-GEMM with split-fp16 operands, fused |.| -> mel -> log)",
                         "bound": "tensor", "achieved": fe_tflops, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": fe_tflops / peaks["tflops"], "peak_source": peaks["source"] + " (sustained bf16)",
                         "algorithmic_flop_per_launch": FLOP_FRONTEND * B, "avg_launch_ms": dft_ms / max(dft_n, 1),
                         "executed_flop_factor": 1.5, "traffic": None,
                         "note": "algorithmic = one K = 480, N = 514 DFT-as-GEMM per clip (SURVEY 8d: 50.7 MFLOP); the kernel "
                                 "takes one radix-2 step out of it (half the products) and executes 3 fp16 products per "
                                 "fp32-accurate product, so the tensor pipe does 1.5x this"},
            "secondary_rooflines": {
                "whole_step_hbm_frac": BYTES_FRONTEND * B * args.steps / (ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                "whole_step_tensor_frac": FLOP_FRONTEND * B * args.steps / (ms / 1e3) / 1e12 / peaks["tflops"],
                "augment_hbm_gbs": BYTES_AUGMENT * B / (aug_ms / 1e3) / 1e9 if aug_ms else None},
            "kernel_classes": {k: {"ms_per_step": v[0] / 5, "launches_per_step": v[1] / 5} for k, v in classes.items() if v[1]},
            "e2e": {"value": world * B * e2e_steps / dt, "unit": "clips/s", "h2d_bytes_per_step": B * 32000 + 5 * B * 4,
                    "d2h_bytes_per_step": B * 15680, "steps": e2e_steps,
                    "api": "kws_pipeline_host_pcm16 (int16 PCM in pinned memory in, log-mel [B,98,40] fp32 out)"},
            "gpu_launches": int(launches), "clocks": clocks})
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = CpuReference((), config=2).rate(seconds_target=10.0, max_batches=64)
        emit(line)
    ctx.finish()
    eng.close()


# -------------------------------------------------------------------------------------------- configs 4 and 5
def run_config45(args):
    from speech_recognition_b200 import Engine, synth, TTA_SHIPPED
    from speech_recognition_b200.classes import class_map_32_to_12
    from speech_recognition_b200.sharded import all_gather_rows
    ctx = Ctx(args)
    torch, dev, world, rank = ctx.torch, ctx.dev, ctx.world, ctx.rank
    views, V = TTA_SHIPPED, 3
    archs = (106,) if args.config == 4 else (106, 195, 206)
    eng = Engine(device=ctx.local, max_rows=args.max_rows, precision=args.precision)
    for slot, a in enumerate(archs):
        eng.load_model(slot, a, synth.synthetic_weights(a))
    cmap_heng, cmap_frozen = class_map_32_to_12("heng"), torch.from_numpy(np.asarray(class_map_32_to_12("frozen"))).to(dev)
    sizes = [args.batch] if (args.config == 4 or args.batch_given) else [1024, 2048, 4096, 8192, 16384, 32768, 65536]
    _, offs = synth.make_noise_bank(seconds=2)
    sweep = []
    main = None
    headline_B = 16384 if 16384 in sizes else sizes[-1]
    for B in sizes:
        x, xi, _ = make_inputs(ctx, B, offs)
        n_all = world * B

        def step():
            if args.config == 4:
                p32, _ = eng.forward(x, views=views, slot=0)
                _, u8 = eng.convert_classes(p32, cmap_heng, 12)
                label, keep = eng.select(u8, 0.6)
                if world > 1:                                     # only the reduced per-clip results cross NVLink
                    return all_gather_rows(u8, n_all), all_gather_rows(label, n_all), all_gather_rows(keep, n_all)
                return u8, label, keep
            labels = []
            for slot, a in enumerate(archs):
                _, am = eng.forward(x, views=views, slot=slot, want_probs=False)
                labels.append(cmap_frozen[am.long()] if a == 106 else am)
            voted, clear = eng.vote(torch.stack(labels).to(torch.int32).contiguous(), 2)
            if world > 1:
                return all_gather_rows(voted, n_all), all_gather_rows(clear, n_all)
            return voted, clear
        steps = args.steps if B >= 8192 else args.steps * 4
        sampler = ClockSampler(ctx.local) if (rank == 0 and B == headline_B) else None
        for _ in range(3):
            step()
        ctx.sync_all()
        if sampler:
            sampler.start()
        l0 = eng.launch_count
        ms = ctx.timed(step, steps, 0)
        launches = eng.launch_count - l0
        clocks = sampler.stop() if sampler else None
        # e2e: int16 PCM in pinned memory -> labels on the host (config 4: uint8 probabilities + label + keep)
        hx16 = pinned(ctx, xi)
        h_p = [torch.empty((B, 32 if a == 106 else 12), dtype=torch.float32).pin_memory() for a in archs]
        h_a = [torch.empty((B,), dtype=torch.int32).pin_memory() for a in archs]

        def e2e():
            for slot, a in enumerate(archs):
                eng.predict_host(hx16.numpy(), views=views, slot=slot, probs_out=h_p[slot].numpy(), argmax_out=h_a[slot].numpy())
            if args.config == 4:
                _, u8 = eng.convert_classes(h_p[0].to(dev, non_blocking=True), cmap_heng, 12)
                label, keep = eng.select(u8, 0.6)
                return u8.cpu(), label.cpu(), keep.cpu()
            labs = torch.stack([cmap_frozen.cpu()[h_a[0].long()].to(torch.int32), h_a[1], h_a[2]]).to(dev)
            voted, clear = eng.vote(labs.contiguous(), 2)
            return voted.cpu(), clear.cpu()
        e_steps = max(1, min(steps, 5))
        dt = ctx.timed_host(e2e, e_steps)
        rec = {"batch_per_gpu": B, "clips_per_s": n_all * steps / (ms / 1e3), "ms_per_step": ms / steps,
               "e2e_clips_per_s": n_all * e_steps / dt, "gpu_launches": int(launches), "steps": steps}
        sweep.append(rec)
        if B == headline_B:
            main = (B, ms, steps, launches, clocks, rec)
        del hx16, x, xi
    if rank == 0:
        B, ms, steps, launches, clocks, rec = main
        peaks = load_peaks()
        flop = sum(FLOP_PER_VIEW[195 if a == 206 else a] for a in archs) * V
        line = base_line(ctx, args, rec["clips_per_s"], ms, steps, {
            "batch_per_gpu": B, "views": V, "models": list(archs),
            "l2_policy": "inputs larger than L2 at B >= 2048 (batch x 64 KB); the 1k point (66 MB) is L2-resident and says so",
            "parallelism": f"dp{world} (clip shards; gathers only uint8 probabilities / labels / flags)"})
        line.update({
            "roofline": {"kernel": "whole step (same tc_gemm / conv1_block1 kernels as config 3)", "bound": "tensor",
                         "achieved": flop * world * B * steps / (ms / 1e3) / 1e12 / world, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": flop * B * steps / (ms / 1e3) / 1e12 / peaks["tflops"], "traffic": None,
                         "peak_source": peaks["source"] + " (sustained bf16)"},
            "sweep": sweep,
            "e2e": {"value": rec["e2e_clips_per_s"], "unit": "clips/s", "h2d_bytes_per_step": len(archs) * B * 32000,
                    "d2h_bytes_per_step": sum(B * (32 if a == 106 else 12) * 4 + B * 4 for a in archs), "steps": max(1, min(steps, 5)),
                    "api": "kws_predict_host_pcm16 per model (ctypes, pinned int16 PCM in) + kws_convert_classes/kws_select or kws_vote"},
            "gpu_launches": int(launches), "clocks": clocks})
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = CpuReference(views, config=args.config).rate(seconds_target=10.0)
        emit(line)
    ctx.finish()
    eng.close()


def _claim_stdout():
    """stdout carries exactly one JSON line (the contract).  Libraries print there too (NCCL's version banner when
    NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for everything else and the
    JSON line is written through a private duplicate of the original stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_JSON_OUT = sys.stdout


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 5], help="BASELINE.json config (1-based); default 3 = the headline")
    ap.add_argument("--job", action="store_true", help="config 3: the fixed 158,538-clip job as one step (strong scaling)")
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU per step (default 16384; config 2: 4096)")
    ap.add_argument("--max-rows", type=int, default=65536, help="clip-views per internal chunk (r01: 32768)")
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--quick", action="store_true", help="skip the e2e variants")
    args = ap.parse_args()
    args.batch_given = args.batch is not None
    if args.batch is None:
        args.batch = 16384
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 3:
        run_config3(args)
    elif args.config == 2:
        run_config2(args)
    else:
        run_config45(args)


if __name__ == "__main__":
    main()
