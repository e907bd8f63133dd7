"""Benchmark of the AudioVidSum hot path on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--axis literal_b1|temporal]

One "step" = score + summarise one batch of BASELINE.json configs[1]: 50 synthetic TVSum-length
videos (T in [200, 700], 21,477 sampled frames, 1024-d visual + 128-d audio features), i.e.
AVBiLSTMModel.forward for every video followed by shot pooling over change points and 0/1
knapsack selection at the 15 % budget.  With N > 1 (torchrun, one process per GPU) the global
batch is 50*N videos sharded by video across the ranks (weak scaling, no data-path collective;
the per-video keyshot picks are gathered with one small NCCL all_gather per step).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` goes through the
public API with pinned HOST buffers (H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import avsum_b200  # noqa: E402,F401
from avsum_b200 import synth, sharding  # noqa: E402

METRIC = "frames/sec scored+summarized"
WORKLOAD = "config2: 50 synthetic TVSum-length videos (T=200-700, 21,477 frames), 1024-d visual + 128-d audio, " \
           "shot pooling + 0/1 knapsack @15%"

# algorithmic FLOPs per frame of each stage, E=1024, H=512, Hc=256
def stage_flops_per_frame(axis: str, mean_T: float):
    f = {
        "fc_gemm": 2 * 512 * (1024 + 128),
        "lstm_input_gemm": 2 * (2 * 512 * 2048),
        "lstm_recurrence": 4 * (2 * 256 * 1024),
        "attn_in_proj_gemm": 2 * 1024 * (3072 if axis == "temporal" else 1024),
        "attention_core": (4 * mean_T * 1024) if axis == "temporal" else 0.0,
        "attn_out_proj_gemm": 2 * 1024 * 1024,
        "score_head_gemm": 2 * (64 * 1024 + 64),
    }
    # fc + LSTM-input GEMMs of both branches, timed as one stage when the audio branch runs on its side stream
    f["frontend_gemms"] = f["fc_gemm"] + f["lstm_input_gemm"]
    return f


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_global_batch(n_gpus: int):
    vids = []
    for c in range(n_gpus):
        vids += synth.video_batch(50, 200, 700, length_seed=c, seed0=1234 + 1000 * c)
    return vids


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU path on the box's host cores: B=1 loop of scripts/evaluate.py:12-18 over
    the torch operators the reference calls (oracle/av_oracle_torch.py, bit-identical port) followed by
    the summary oracle.  /root/reference does not exist on the GPU box and is not imported."""
    if rank != 0:
        return
    from oracle import av_oracle, av_oracle_torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vids = synth.config2()
    n_sample = len(vids) if (args.steps + args.warmup) <= 16 else 10
    sample = vids[:n_sample]
    model = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    model.load_state_dict(synth.seeded_state_dict())
    frames = sum(v.T for v in sample)

    def step():
        scores = av_oracle_torch.run_videos(model, [(v.visual, v.audio) for v in sample], "literal")
        for v, s in zip(sample, scores):
            av_oracle.generate_summary(s.numpy(), v.cps, v.n_frames, v.positions)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = frames / dt
    desc = f"{n_sample} of the 50 config-2 videos ({frames} frames) per step, B=1 loop, torch CPU ops + numpy summary"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "attn_axis": "literal (B=1 per video)"},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "videos_per_s": n_sample / dt,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--axis", default="literal_b1", choices=["literal_b1", "temporal"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from avsum_b200 import _cabi
    from avsum_b200.models.av_model import AVBiLSTMModel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: keep the pinned batches of this rank in the memory next to its GPU (no-op on one node)
    print(f"[rank {rank}] " + sharding.bind_process_to_gpu_numa(local_rank), file=sys.stderr, flush=True)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION/INFO
        # (also at WARN); NCCL logging is therefore off unless AVS_NCCL_DEBUG asks for it, and then goes to a file
        os.environ.pop("NCCL_DEBUG", None)
        if os.environ.get("AVS_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = os.environ["AVS_NCCL_DEBUG"]
            os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/avs_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload: global batch sharded by video
    vids_all = build_global_batch(max(world, 1))
    shards = sharding.shard_videos([v.T for v in vids_all], world)
    mine = shards[rank]
    # batch composition as data/dataset.py packed_batches builds it: longest video first (the host-space call
    # pipelines the batch by video group, and a group's recurrence lasts as long as its longest video)
    mine = sorted(mine, key=lambda i: -vids_all[i].T)
    vids = [vids_all[i] for i in mine]
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    R = int(sum(lens))
    frames_global = sum(v.T for v in vids_all)

    model = AVBiLSTMModel(1024, 128, 512, attn_axis=args.axis).eval()
    model.load_state_dict(synth.seeded_state_dict())
    model = model.to(dev)
    nat = model.native()

    # pinned host batch, as a loader packs it.  AVS_BENCH_WC=1: write-combined pages (runtime.pinned_like)
    from avsum_b200 import runtime as _rt
    if os.environ.get("AVS_BENCH_WC"):
        visual_h = _rt.pinned_like(torch.cat([v.visual for v in vids]), write_combined=True)
        audio_h = _rt.pinned_like(torch.cat([v.audio for v in vids]), write_combined=True)
    else:
        visual_h = torch.cat([v.visual for v in vids]).pin_memory()
        audio_h = torch.cat([v.audio for v in vids]).pin_memory()
    pos_h = torch.from_numpy(np.concatenate([v.positions for v in vids]).astype(np.int32)).pin_memory()
    visual_d, audio_d, pos_d = torch.cat([v.visual for v in vids]).to(dev), torch.cat([v.audio for v in vids]).to(dev), pos_h.to(dev)
    n_frames = [v.n_frames for v in vids]
    cps_list = [v.cps for v in vids]
    from avsum_b200.evaluation.summary import summarize_stream
    from avsum_b200.runtime import ShotDesc
    shots = ShotDesc(n_frames, cps_list)     # packed once with the batch, like row_start / lengths
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # result gather (the only collective of the inference path): every rank knows every shard's shot count
    # from the host-side change points, so one padded all_gather of the keyshot picks per step suffices
    shots_per_rank = [sum(len(vids_all[i].cps) for i in sh) for sh in shards]
    pad = torch.zeros(max(shots_per_rank), dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * pad.numel(), dtype=torch.uint8, device=dev)

    def step_device():
        scores = nat.forward_rows(visual_d, audio_d, starts, lens, args.axis, "tf32")
        picks, seg_mean, summary, cps_start, _ = nat.summarize_rows(scores, pos_d, starts, lens, None, shots, 0.15)
        if world > 1:
            pad[:picks.numel()].copy_(picks)
            dist.all_gather_into_tensor(gathered, pad)
        return picks

    def step_host():
        # the public "score + summarise" call with pinned HOST buffers: features cross PCIe inside the call
        # (pipelined by video group), scores / picks / shot means / keyshot bitmap come back to host memory
        scores, picks, seg_mean, summary, _, _ = nat.score_and_summarize_rows(
            visual_h, audio_h, pos_h, starts, lens, None, shots, 0.15, args.axis, "tf32")
        return scores, picks, seg_mean, summary

    def stream_host(k):
        # the same call as a dataset loop makes it (scripts/evaluate.py:12-18 over packed batches): k batches through
        # evaluation.summary.summarize_stream, two in flight, so batch i+1 crosses PCIe while batch i is computed;
        # every batch's features go host -> device and its scores / picks / shot means / bitmap come back
        last = None
        for last in summarize_stream(model, ((visual_h, audio_h, pos_h, starts, lens, shots) for _ in range(k)),
                                     0.15, args.axis):
            pass
        return last[:4]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()

    # ---- timed region (device-resident inputs), CUDA events per step, L2 flushed between steps
    _cabi.profile(2)
    launches0 = _cabi.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1)
        a.record()
        step_device()
        b.record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = _cabi.launch_count() - launches0
    stage_ms = _cabi.profile_read()
    _cabi.profile(0)
    ms_local = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([ms_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item())

    # ---- end-to-end through the public API with pinned host buffers.
    # (1) one synchronous call per batch (latency of a single "score + summarise" call), L2 flushed before each
    for _ in range(2):
        step_host()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        res = step_host()
    barrier()
    call_wall = time.perf_counter() - e0
    torch.cuda.synchronize()
    f0 = time.perf_counter()
    for _ in range(args.steps):       # the flush is not part of the step: time it alone and subtract
        flush.fill_(1)
        torch.cuda.synchronize()
    flush_wall = time.perf_counter() - f0
    call_ms_local = (call_wall - flush_wall) / args.steps * 1e3
    # (2) the streamed form: K batches back to back, two in flight.  Every step moves its 99 MB of features
    # through one of two alternating device staging areas and ~0.6 GB of activations -- far more than the 126 MB
    # L2 -- so no explicit flush is interleaved.
    stream_host(4)
    barrier()
    e0 = time.perf_counter()
    res = stream_host(args.steps)
    barrier()
    e2e_ms_local = (time.perf_counter() - e0) / args.steps * 1e3
    # clocks / throttle reasons sampled over ALL timed regions (device-resident steps, single calls, streamed steps):
    # the device-resident region alone lasts ~20 ms, less than two sampling periods
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([e2e_ms_local, call_ms_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, call_ms = float(t[0].item()), float(t[1].item())
    scores_h, picks_h, segm_h, summ_h = res
    h2d = R * (1024 + 128) * 4 + R * 4           # features + frame positions
    d2h = R * 4 + picks_h.numel() + segm_h.numel() * 8 + summ_h.numel()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events inside the library, same timed region)
    peaks = load_peaks()
    mean_T = float(np.mean([t * t for t in lens]) / np.mean(lens))  # frame-weighted mean length
    flops = stage_flops_per_frame(args.axis, mean_T)
    kernels = {}
    for name, (ms, calls) in stage_ms.items():
        if calls == 0:
            continue
        per = ms / args.steps
        rec = {"ms_per_step": per}
        if name in flops and flops[name] > 0:
            rec["tflops"] = flops[name] * R / (per * 1e-3) / 1e12
            rec["frac_of_tensor_peak"] = rec["tflops"] / peaks["tflops"]
        if name == "shot_pool":
            nbytes = 8 * R + 16 * sum(len(c) for c in cps_list)
            rec["gbs"] = nbytes / (per * 1e-3) / 1e9
            rec["frac_of_hbm_peak"] = rec["gbs"] / peaks["hbm_gbs"]
        if name == "convert_tf32":
            nbytes = 2 * R * (1024 + 128) * 4
            rec["gbs"] = nbytes / (per * 1e-3) / 1e9
            rec["frac_of_hbm_peak"] = rec["gbs"] / peaks["hbm_gbs"]
        kernels[name] = rec
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    drec = kernels[dom]
    # DRAM traffic per launch of the dominant kernel, from the committed ncu --set full capture of this workload
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01d_traffic.json")
    if world == 1 and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)
    if "tflops" in drec:
        roofline = {"kernel": dom, "bound": "tensor", "achieved": drec["tflops"], "peak": peaks["tflops"], "unit": "TFLOP/s",
                    "frac": drec["tflops"] / peaks["tflops"], "traffic": traffic,
                    "peak_source": peaks["source"] + " bf16 sustained (MEASURED_PEAKS.json)",
                    "share_of_step": drec["ms_per_step"] / ms_per_step,
                    "note": "the LSTM recurrence is a chain of max(T) dependent steps (latency bound, SURVEY 8d); "
                            "see kernels{} for the GEMM / attention tensor-pipe fractions"}
    else:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": drec.get("gbs"), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": (drec.get("gbs") or 0) / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["source"],
                    "share_of_step": drec["ms_per_step"] / ms_per_step}

    line = {
        "metric": METRIC, "value": frames_global / (ms_per_step * 1e-3), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "tf32 features / fp16 activations (11-bit significands), fp32 accumulate, state and scores",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "videos_per_gpu": len(vids), "frames_per_gpu": R, "global_videos": len(vids_all),
                   "attn_axis": args.axis, "l2": "256 MiB flush between timed steps", "batch_order": "longest video first (packed_batches)", "parallelism": f"dp{world} by video"},
        "videos_per_s": len(vids_all) / (ms_per_step * 1e-3),
        "e2e": {"value": frames_global / (e2e_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "mode": "streamed: evaluation.summary.summarize_stream, two batches in flight "
                        "(avs_forward_summarize_async); every step's H2D and D2H inside the timed region",
                "single_call_ms": call_ms, "single_call_value": frames_global / (call_ms * 1e-3),
                "l2": "per step 99 MB of features through alternating staging slots + 0.6 GB of activations >> L2; "
                      "single_call: 256 MiB flush before every call"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
        "wall_s_timed_region": wall,
    }

    # ---- CPU baseline (reference CPU path port) on this box's host cores, N=1 only
    if world == 1 and not args.no_cpu_baseline:
        from oracle import av_oracle, av_oracle_torch
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
        port.load_state_dict(synth.seeded_state_dict())
        base_vids = synth.config2()

        def cpu_step():
            sc = av_oracle_torch.run_videos(port, [(v.visual, v.audio) for v in base_vids], "literal")
            for v, s in zip(base_vids, sc):
                av_oracle.generate_summary(s.numpy(), v.cps, v.n_frames, v.positions)

        cpu_step()
        reps, c0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - c0 < 10.0 and reps < 20):
            cpu_step()
            reps += 1
        cdt = (time.perf_counter() - c0) / reps
        line["cpu_baseline"] = {"value": 21477 / cdt, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"the full 50-video config-2 batch x {reps} repetitions, B=1 loop "
                                          "(scripts/evaluate.py:12-18) over torch CPU ops + numpy summary oracle"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
