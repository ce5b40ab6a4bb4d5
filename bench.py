#!/usr/bin/env python
"""Benchmark of the keyword-spotting hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference path

Workload (config.workload): BASELINE.json configs[2] = exp-195 Depthwise1D forward with 8x TTA, with
the north-star stages in front of it -- one "step" is one pass over a batch of B synthetic 1 s / 16 kHz
clips: augment (time-shift + noise mix + volumes) -> log-mel (40 mel, 30 ms / 10 ms) -> 8-view forward
-> TTA mean -> argmax.  B clips x 64 KB = 1.05 GB at the default B=16384 (about the per-GPU share of
the 158,538-clip job on 8 GPUs), far larger than the 126 MB L2; the forward runs in chunks of
--max-rows clip-views (32768 = 4096 clips).
Per-GPU work is fixed as N grows (weak scaling); the only collective is one all-gather of the
[B,12] probabilities per step.  `value` times the device-resident path with CUDA events on the
launching stream (max over ranks); `e2e` times the host-buffer C-ABI call (pinned host memory,
H2D + D2H inside) with the same work.
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
FLOP_PER_VIEW = 112.48e6                   # SURVEY.md 8(d), exp 195/206
FLOP_CONV1, FLOP_HEAD = 12.26e6, 0.11e6
FLOP_BLOCKS = FLOP_PER_VIEW - FLOP_CONV1 - FLOP_HEAD
FLOP_FRONTEND = 50.7e6                     # DFT-as-GEMM + mel + DCT per clip
BYTES_AUGMENT = 192020                     # per clip
ACT_BYTES_VIEW_FP16 = 286720 * 2 * 2       # 12 block outputs written + read once, fp16
WORKLOAD = ("config3: augment + log-mel(40 mel, 30/10 ms) + exp-195 Depthwise1D forward x 8 TTA views + TTA mean "
            "+ argmax; one step = one batch")


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


# --------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle), timed on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_rate(views, seconds_target=20.0, batch=64, max_batches=16):
    """clips/s of augment + log-mel + n-view forward with the oracle on all host threads.
    bench.py may execute oracle/ only here (cpu_baseline / --impl reference)."""
    import torch
    from oracle import augment, frontend, network, driver
    from speech_recognition_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    clips = synth.make_clips(batch, seed=synth.SEED + 3)
    bank, offs = synth.make_noise_bank(seconds=4)
    p = synth.make_params(batch, offs, seed=synth.SEED + 4)
    w = synth.synthetic_weights(195)

    def one_batch():
        bg = augment.gather_background(bank, offs, p["bg_index"], p["bg_offset"])
        x = augment.augment_mix(clips, p["time_shift"], bg, p["bg_volume"], p["fg_volume"])
        frontend.features(x, kind="logmel", dct_coefficient_count=40, fft_dtype=np.float32)
        return driver.tta_predict(lambda v: network.forward(v, w, 195, dtype=torch.float32), x, views)
    one_batch()                                              # warm-up (oneDNN primitive caches)
    t0 = time.perf_counter(); one_batch(); t1 = time.perf_counter() - t0
    n = int(max(1, min(max_batches, seconds_target / max(t1, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        one_batch()
    dt = time.perf_counter() - t0
    return dict(value=batch * n / dt, unit="clips/s", cores=threads, kind="port",
                sample=f"{n} batches of {batch} clips x {len(views)} views (aug+logmel+fwd), torch-CPU fp32 oracle")


def run_reference(args):
    from speech_recognition_b200 import TTA_8
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    res = []
    for _ in range(max(1, min(args.steps, 3))):
        res.append(cpu_reference_rate(TTA_8, seconds_target=args.ref_seconds))
    best = max(res, key=lambda r: r["value"])
    line = {
        "impl": "reference", "metric": "1s-clips/sec (aug+feat+fwd, 8x TTA)", "value": best["value"],
        "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * 64 / best["value"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "views": 8, "job_clips": N_CLIPS_JOB,
                   "reference_arm": "CPU restatement of the reference path (oracle/, torch-CPU fp32, all host threads; "
                                    "TF 1.4 / Keras 2.1.2 are not installable here), bounded sample of the workload: "
                                    "batches of 64 clips x 8 views"},
        "cpu_baseline": best,
        "e2e": {"value": best["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from speech_recognition_b200 import Engine, synth, TTA_8

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    B, V = args.batch, 8
    views = TTA_8
    eng = Engine(device=local, max_rows=args.max_rows, precision=args.precision)
    bank, offs = synth.make_noise_bank(seconds=60)                 # 23 MB, resident per GPU
    eng.set_noise_bank(torch.from_numpy(bank).to(dev), offs)
    eng.frontend_config(480, 160, 40, 40)                          # BASELINE config 2 front end
    eng.load_model(0, 195, synth.synthetic_weights(195))
    # synthetic clips: a pool of distinct clips tiled to the batch (every rank a different slice)
    pool = torch.from_numpy(synth.make_clips(512, seed=synth.SEED + 17 * rank)).to(dev)
    x = pool.repeat((B + 511) // 512, 1)[:B].contiguous()
    p = synth.make_params(B, offs, seed=synth.SEED + 1 + rank)
    pt = {k: torch.from_numpy(v).to(dev) for k, v in p.items()}
    aug = torch.empty_like(x)
    feat = torch.empty((B, 98, 40), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * B, 12), dtype=torch.float32, device=dev) if world > 1 else None

    def step():
        eng.augment(x, pt["time_shift"], pt["bg_index"], pt["bg_offset"], pt["bg_volume"], pt["fg_volume"], out_t=aug)
        eng.features(aug, "logmel", out_t=feat)
        probs, amax = eng.forward(aug, views=views)
        if world > 1:
            dist.all_gather_into_tensor(gathered, probs)
        return probs, amax

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    eng.timing_read()
    eng.timing_enable(True)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        probs, amax = step()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count - launches0
    classes = eng.timing_read()
    eng.timing_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- e2e: host buffers (pinned) through the C-ABI host entry point ----
    hx = torch.empty((B, 16000), dtype=torch.float32).pin_memory()
    hx.copy_(x.cpu())
    hp = {}
    for k, v in p.items():
        tpin = torch.from_numpy(v).pin_memory()
        hp[k] = tpin.numpy()
    h_feat = torch.empty((B, 98 * 40), dtype=torch.float32).pin_memory()
    h_probs = torch.empty((B, 12), dtype=torch.float32).pin_memory()
    h_amax = torch.empty((B,), dtype=torch.int32).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        eng.pipeline_host(hx.numpy(), hp, feat_kind="logmel", views=views, feat_out=h_feat.numpy(),
                          probs_out=h_probs.numpy(), argmax_out=h_amax.numpy())
    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t.item())
    h2d = B * 16000 * 4 + 5 * B * 4
    d2h = B * 98 * 40 * 4 + B * 12 * 4 + B * 4

    if rank == 0:
        peaks = load_peaks()
        blk_ms, blk_n = classes["dw_pw_blocks"]
        views_per_step = B * V
        blk_tflops = FLOP_BLOCKS * views_per_step * args.steps / (blk_ms / 1e3) / 1e12 if blk_ms > 0 else 0.0
        # dominant kernel = the 11 block launches of each chunk; roofline on its algorithmic HBM bytes
        bpv = block_bytes_per_view(195)
        views_per_launch = min(args.max_rows // V * V, views_per_step)
        per_block = getattr(eng, "last_block_ms", None)
        # blocks that ran as tc_gemm_kernel launches (block 1 is fused into the conv1d_1 kernel when shapes allow)
        active = [i for i in range(len(bpv)) if per_block and per_block[i][1] > 0] or list(range(len(bpv)))
        alg_bytes_per_launch = sum(bpv[i] for i in active) / len(active) * views_per_launch
        blk_avg_launch_ms = blk_ms / max(blk_n, 1)
        blk_gbs = alg_bytes_per_launch / (blk_avg_launch_ms / 1e3) / 1e9 if blk_ms > 0 else 0.0
        blk_flop = FLOP_BLOCKS * sum(bpv[i] for i in active) / sum(bpv)      # tensor view: approximate share of the active blocks
        blk_tflops = blk_flop * views_per_step * args.steps / (blk_ms / 1e3) / 1e12 if blk_ms > 0 else 0.0
        fused_first = bool(per_block) and per_block[0][1] == 0
        tr = measured_traffic()
        per_class = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
                     for k, v in classes.items() if v[1]}
        aug_ms = classes["augment"][0] / args.steps
        line = {
            "metric": "1s-clips/sec (aug+feat+fwd, 8x TTA)", "value": value, "unit": "clips/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if args.precision == "tc" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": B, "views": V, "job_clips": N_CLIPS_JOB,
                       "l2_policy": "inputs larger than L2 (batch x 64 KB = %.0f MB)" % (B * 64e3 / 1e6),
                       "precision": args.precision, "max_rows": args.max_rows,
                       "fused_conv1_block1": fused_first,
                       "parallelism": f"dp{world} (clip shards, 1 all-gather of probabilities per step)"},
            "roofline": {"kernel": ("tc_gemm_kernel<1|2> (TMA-fed depthwise producer + tcgen05 pointwise GEMM + BN/ReLU6 "
                                    "+ TMA store), %d launches per chunk" % len(active)) if args.precision == "tc" else "gemm_f32_kernel",
                         "bound": "hbm", "achieved": blk_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": blk_gbs / peaks["hbm_gbs"], "peak_source": peaks["source"] + " (copy bandwidth)",
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                         "avg_launch_ms": blk_avg_launch_ms, "clip_views_per_launch": views_per_launch,
                         "traffic": tr["bytes_per_launch"] if tr else None,
                         "traffic_source": tr["source"] if tr else None,
                         "share_of_step": blk_ms / ms if ms else None,
                         "tensor_tflops": blk_tflops, "tensor_frac": blk_tflops / peaks["tflops"]},
            "secondary_rooflines": {
                "augment_hbm_gbs": BYTES_AUGMENT * B / (aug_ms / 1e3) / 1e9 if aug_ms else None,
                "augment_hbm_frac": (BYTES_AUGMENT * B / (aug_ms / 1e3) / 1e9) / peaks["hbm_gbs"] if aug_ms else None,
                "whole_step_tensor_frac": (FLOP_PER_VIEW * V + FLOP_FRONTEND) * B * args.steps / (ms / 1e3) / 1e12 / peaks["tflops"],
                "whole_step_hbm_frac_block_materialised": (ACT_BYTES_VIEW_FP16 * V + BYTES_AUGMENT + 64000 + 15680) * B
                * args.steps / (ms / 1e3) / 1e9 / peaks["hbm_gbs"],
            },
            "kernel_classes": per_class,
            "block_ms_per_step": [round(b[0] / args.steps, 4) for b in per_block] if per_block else None,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "kws_pipeline_host (ctypes, pinned host buffers)"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_rate(views, seconds_target=15.0)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
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
    ap.add_argument("--batch", type=int, default=16384, help="clips per GPU per step")
    ap.add_argument("--max-rows", type=int, default=32768, help="clip-views per internal chunk")
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-seconds", type=float, default=15.0, help="--impl reference: CPU seconds per step (bounded sample)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
