"""Benchmark of the AudioVidSum hot path on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config infer|long|train] [--axis literal_b1|temporal]

--config infer (default; BASELINE.json configs[1], the configuration the metric is quoted on).  One "step" = score +
    summarise one batch of 50 synthetic TVSum-length videos (T in [200, 700], 21,477 sampled frames, 1024-d visual +
    128-d audio features): AVBiLSTMModel.forward for every video, shot pooling over change points, 0/1 knapsack at
    the 15 % budget.  With N > 1 (torchrun, one process per GPU) the global batch is 50*N videos sharded by video
    (weak scaling, no data-path collective; the keyshot picks are gathered with one small NCCL all_gather per step).
    The same JSON line carries the second half of the metric, `attention`: a configs[3]-shaped run (8 videos x
    T = 8192, temporal attention) with the tcgen05 attention core's time, TFLOP/s and fraction of the measured bf16
    peak, and `parity`: the timed step's scores / keyshots against the CPU port of the reference.
--config long   BASELINE.json configs[3]: 8*N videos x T = 8192, temporal attention, sharded by video.
--config train  BASELINE.json configs[4]: training step (8 videos x T = 320 per GPU, train mode, mse_loss, backward,
    bucketed NCCL gradient all-reduce overlapped with the backward, AdamW).

Prints ONE JSON line (rank 0) on stdout.  `value` is device-resident throughput; `e2e` goes through the public API
with pinned HOST buffers (H2D + D2H inside the timed region).  NCCL_DEBUG is left as the caller set it; when it is
set and NCCL_DEBUG_FILE is not, NCCL's log lines go to stderr so that stdout stays the one JSON line.
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
WORKLOADS = {
    "infer": "config2: 50 synthetic TVSum-length videos (T=200-700, 21,477 frames), 1024-d visual + 128-d audio, "
             "shot pooling + 0/1 knapsack @15%",
    "long": "config4: 8 synthetic videos x T=8192 frames per GPU, 1024-d visual + 128-d audio, temporal attention, "
            "shot pooling + 0/1 knapsack @15%",
    "train": "config5: training step, 8 synthetic videos x T=320 per GPU, train mode (dropout 0.3), mse_loss, "
             "backward, NCCL gradient all-reduce (32 MB fp32, bucketed, overlapped), AdamW lr 1e-4",
}
DTYPE = "tf32 features / fp16 activations (11-bit significands), fp32 accumulate, state and scores"
TRAFFIC_PROFILE = os.path.join("profiles", "r02_traffic.json")


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
    # out_proj with the score head fused into its epilogue
    f["out_proj_score_gemm"] = f["attn_out_proj_gemm"] + f["score_head_gemm"]
    # pipelined front: fc + LSTM-input GEMMs of group k+1.. run beside the recurrence of groups ..k; one timed stage
    f["frontend_lstm_pipelined"] = f["frontend_gemms"] + f["lstm_recurrence"]
    return f


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tflops_burst": d["bf16_tflops"], "source": "measured"}
    # fallback stated by /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_burst": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed regions."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
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


def global_batch_spec(n_gpus: int, config: str = "infer"):
    """(T, seed) of every video of the global batch: 50 TVSum-length videos (config 2) or 8 x T=8192 (config 4) per
    GPU.  Cheap -- every rank needs all lengths / shot counts for the sharding, but features only for its shard."""
    spec = []
    for c in range(n_gpus):
        if config == "long":
            spec += [(8192, 9000 + 100 * c + i) for i in range(8)]
        else:
            lengths = np.random.default_rng(c).integers(200, 701, 50)
            spec += [(int(t), 1234 + 1000 * c + i) for i, t in enumerate(lengths)]
    return spec


def n_shots(T: int, seed: int) -> int:
    return int(synth.make_change_points(synth.SAMPLE_STRIDE * T, seed=100000 + seed).shape[0])


def build_global_batch(n_gpus: int, config: str = "infer"):
    return [synth.make_video(T, 1024, 128, seed) for T, seed in global_batch_spec(n_gpus, config)]


class _StdoutToStderr:
    """NCCL prints its version banner with a bare printf to stdout when a communicator is created (NCCL_DEBUG=VERSION
    or INFO); stdout must hold only the JSON line, so file descriptor 1 points at stderr while the process group is
    set up -- the banner and every NCCL_DEBUG line stay visible to whoever reads stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def finish(world, dist):
    """Leave without tearing NCCL down: destroy_process_group() after a CUDA graph with captured collectives was
    observed to hang (2 x B200, r02e); every rank has synchronised and rank 0 has printed, so exit directly."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU path on the box's host cores: B=1 loop of scripts/evaluate.py:12-18 over the torch
    operators the reference calls (oracle/av_oracle_torch.py, bit-identical port) followed by the summary oracle, on
    the SAME workload as our arm's N=1 config (all 50 config-2 videos per step; 8 x T=8192 for --config long; the
    8 x 320 training step for --config train).  /root/reference does not exist on the GPU box and is not imported."""
    if rank != 0:
        return
    from oracle import av_oracle, av_oracle_torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = args.config
    if cfg == "train":
        return run_reference_train(args, cores)
    sample = build_global_batch(1, cfg)
    axis = "temporal" if (cfg == "long" or args.axis == "temporal") else "literal"
    model = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    model.load_state_dict(synth.seeded_state_dict())
    frames = sum(v.T for v in sample)

    def step():
        scores = av_oracle_torch.run_videos(model, [(v.visual, v.audio) for v in sample], axis)
        for v, s in zip(sample, scores):
            av_oracle.generate_summary(s.numpy(), v.cps, v.n_frames, v.positions)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = frames / dt
    desc = f"all {len(sample)} videos of the workload ({frames} frames) per step, B=1 loop, torch CPU ops + numpy summary"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg], "attn_axis": axis + " (B=1 per video)", "videos_per_step": len(sample),
                   "frames_per_step": frames},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "videos_per_s": len(sample) / dt,
    }
    print(json.dumps(line), flush=True)


def run_reference_train(args, cores):
    from oracle import av_oracle_torch
    B, T = 8, 320
    g = torch.Generator().manual_seed(100)
    visual, audio = torch.randn(B, T, 1024, generator=g), torch.randn(B, T, 128, generator=g)
    target = torch.rand(B, T, generator=g)
    port = av_oracle_torch.RefPortModel(1024, 128, 512).train()
    port.load_state_dict(synth.seeded_state_dict())
    opt = torch.optim.AdamW(port.parameters(), lr=1e-4)

    def step():
        opt.zero_grad()
        for b in range(B):      # the reference's loop: one video per forward/backward (train_av_model.py:86-96)
            loss = torch.nn.functional.mse_loss(port(visual[b:b + 1], audio[b:b + 1]), target[b]) / B
            loss.backward()
        opt.step()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = B * T / dt
    line = {"impl": "reference", "metric": "frames/sec trained", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS["train"]},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "the full 8 x 320 training step, torch CPU autograd + AdamW"},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ helpers of our arm
def setup_dist(world, dev):
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG stays as the caller (the driver) set it, so that the run proves its rank count; NCCL writes its
        # log to stdout by default, which must hold only the JSON line -> send it to stderr unless told otherwise
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        with _StdoutToStderr():
            dist.init_process_group("nccl", device_id=dev)
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)            # the communicator (and NCCL's banner) exist before anything is printed
            torch.cuda.synchronize()
    return dist


def max_over_ranks(values, dev, world, dist):
    t = torch.tensor(list(values), device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def h2d_floor(host_tensors, dev, world, dist, trials=3, reps=4):
    """Bare pinned host -> device copy of one step's inputs, all ranks at once: the floor of the end-to-end step on
    this box (PCIe / host memory).  The copies read the SAME pinned buffers the end-to-end step reads (a separate
    allocation can sit on other pages / another NUMA node) and are issued three ways -- one cudaMemcpyAsync per
    buffer, and 16 MB / 4 MB pieces (the library copies a batch as ~10 pieces, one per video group and modality;
    whole-buffer copies have measured up to 5 % slower than that on boxes of this pool).  Best of all trials and
    piece sizes (a floor is the best the box can do; other tenants of the host show up as slower trials), max over
    ranks.  A GPU busy with compute on another stream copies no faster (tried: 1.925 vs 1.928 ms)."""
    srcs = [t.reshape(-1).view(torch.uint8) for t in host_tensors]
    dsts = [torch.empty(t.numel(), dtype=torch.uint8, device=dev) for t in srcs]
    for _ in range(2):
        for d, t in zip(dsts, srcs):
            d.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    best = float("inf")
    by_piece = {}
    side = torch.cuda.Stream(device=dev)      # the library copies on a non-blocking stream of its own, not on stream 0
    for piece, on_side in ((0, False), (16 << 20, False), (4 << 20, False), (0, True), (16 << 20, True)):
        pairs = []
        for d, t in zip(dsts, srcs):
            n = t.numel()
            step = n if piece == 0 else piece
            pairs += [(d[o:o + step], t[o:o + step]) for o in range(0, n, step)]
        stream = side if on_side else torch.cuda.current_stream(dev)
        for _ in range(trials):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record()
                for _ in range(reps):
                    for d, t in pairs:
                        d.copy_(t, non_blocking=True)
                e1.record()
            torch.cuda.synchronize()
            ms = max_over_ranks([e0.elapsed_time(e1) / reps], dev, world, dist)[0]
            best = min(best, ms)
            key = ("whole_buffers" if piece == 0 else f"{piece >> 20}MB_pieces") + ("_side_stream" if on_side else "")
            by_piece[key] = min(by_piece.get(key, float("inf")), ms)
    h2d_floor.last = by_piece   # best time per piece size of the last probe (reported beside the floor)
    return best


def kernel_table(stage_ms, steps, flops, rows, peaks, pool_bytes=None):
    kernels = {}
    for name, (ms, calls) in stage_ms.items():
        if calls == 0:
            continue
        per = ms / steps
        rec = {"ms_per_step": per}
        if name in flops and flops[name] > 0:
            rec["tflops"] = flops[name] * rows / (per * 1e-3) / 1e12
            rec["frac_of_tensor_peak"] = rec["tflops"] / peaks["tflops"]
        if name == "shot_pool" and pool_bytes:
            rec["gbs"] = pool_bytes / (per * 1e-3) / 1e9
            rec["frac_of_hbm_peak"] = rec["gbs"] / peaks["hbm_gbs"]
        if name == "convert_tf32":
            nbytes = 2 * rows * (1024 + 128) * 4
            rec["gbs"] = nbytes / (per * 1e-3) / 1e9
            rec["frac_of_hbm_peak"] = rec["gbs"] / peaks["hbm_gbs"]
        kernels[name] = rec
    return kernels


def roofline_of(kernels, ms_per_step, peaks, world):
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    drec = kernels[dom]
    # DRAM traffic per launch of the dominant kernel comes from this round's committed ncu --set full capture of the
    # same workload; it is used only when that capture lists the kernel that is dominant in THIS run
    traffic, tsrc = None, None
    tpath = os.path.join(ROOT, TRAFFIC_PROFILE)
    if world == 1 and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if dom in tj:
            traffic = tj[dom]
            tsrc = f"{TRAFFIC_PROFILE} ({tj.get('_source', 'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum')})"
    if "tflops" in drec:
        roof = {"kernel": dom, "bound": "tensor", "achieved": drec["tflops"], "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": drec["tflops"] / peaks["tflops"], "traffic": traffic, "traffic_source": tsrc,
                "peak_source": peaks["source"] + " bf16 sustained (MEASURED_PEAKS.json)",
                "share_of_step": drec["ms_per_step"] / ms_per_step,
                "note": "the LSTM recurrence is a chain of max(T) dependent steps (latency bound, SURVEY 8d); "
                        "see kernels{} and attention{} for the GEMM / attention tensor-pipe fractions"}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": drec.get("gbs"), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": (drec.get("gbs") or 0) / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": tsrc,
                "peak_source": peaks["source"], "share_of_step": drec["ms_per_step"] / ms_per_step}
    return roof


def attention_probe(model_sd, dev, peaks, world, dist, n_videos=8, T=8192, steps=4):
    """BASELINE.json's metric, second half: the tcgen05 attention core on a configs[3]-shaped batch (n_videos x
    T = 8192, temporal attention), timed with CUDA events on the launching stream inside the library."""
    from avsum_b200 import _cabi
    from avsum_b200.models.av_model import AVBiLSTMModel
    model = AVBiLSTMModel(1024, 128, 512, attn_axis="temporal").eval()
    model.load_state_dict(model_sd)
    model = model.to(dev)
    nat = model.native()
    g = torch.Generator(device=dev).manual_seed(4)
    visual = torch.randn(n_videos * T, 1024, generator=g, device=dev)
    audio = torch.randn(n_videos * T, 128, generator=g, device=dev)
    lens = [T] * n_videos
    starts = [i * T for i in range(n_videos)]
    for _ in range(2):
        nat.forward_rows(visual, audio, starts, lens, "temporal", "tf32")
    torch.cuda.synchronize()

    def timed_forwards():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            nat.forward_rows(visual, audio, starts, lens, "temporal", "tf32")
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    fwd_ms = timed_forwards()   # the default schedule (pipelined tail: per video group)
    # the attention core as ONE launch over all videos, timed by the library's stage events: the one-launch schedule
    os.environ["AVS_PIPE_TAIL"] = "0"
    try:
        nat.forward_rows(visual, audio, starts, lens, "temporal", "tf32")
        torch.cuda.synchronize()
        _cabi.profile(2)
        fwd_one_ms = timed_forwards()
        st = _cabi.profile_read()
        _cabi.profile(0)
    finally:
        os.environ.pop("AVS_PIPE_TAIL", None)
    att_ms = st["attention_core"][0] / steps
    att_ms, fwd_ms, fwd_one_ms = max_over_ranks([att_ms, fwd_ms, fwd_one_ms], dev, world, dist)
    flops = 4.0 * T * T * 1024 * n_videos
    tf = flops / (att_ms * 1e-3) / 1e12
    del model, nat, visual, audio
    torch.cuda.empty_cache()
    return {"workload": f"config4-shaped: {n_videos} videos x T={T} per GPU, temporal attention, 4 heads x 256",
            "kernel": "attention_tc_kernel (tcgen05 QK^T / PV, TMEM-resident S / P / O)",
            "ms": att_ms, "tflops": tf, "frac_of_sustained_bf16_peak": tf / peaks["tflops"],
            "frac_of_burst_bf16_peak": tf / peaks["tflops_burst"], "algorithmic_flop": flops,
            "full_forward_ms": fwd_ms, "full_forward_frames_per_s": n_videos * T / (fwd_ms * 1e-3),
            "full_forward_one_launch_schedule_ms": fwd_one_ms,
            "stages_ms": {k: v[0] / steps for k, v in st.items() if v[1]},
            "stages_note": "stage times (and the attention launch) are measured in the one-launch schedule "
                           "(AVS_PIPE_TAIL=0); full_forward_ms is the default schedule, which runs each video group's "
                           "tail behind its own recurrence",
            "ncu": "profiles/r02p_ncu_full_summary.csv, row attention_tc_kernel (ncu --set full of tools/prof_long.py 8 8192 2: "
                   "sm__pipe_tensor_cycles_active 65.0 %, dram__bytes 518.6 MB per launch)"}


# ------------------------------------------------------------------------------------ our arm: inference configs
def run_infer(args, rank, world, local_rank):
    from avsum_b200 import _cabi
    from avsum_b200.models.av_model import AVBiLSTMModel
    from avsum_b200.evaluation.summary import summarize_stream
    from avsum_b200.runtime import ShotDesc
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: keep the pinned batches of this rank in the memory next to its GPU (no-op on one node)
    print(f"[rank {rank}] " + sharding.bind_process_to_gpu_numa(local_rank), file=sys.stderr, flush=True)
    dist = setup_dist(world, dev)
    cfg = args.config
    axis = "temporal" if cfg == "long" else args.axis

    # ---- workload: global batch sharded by video
    spec = global_batch_spec(max(world, 1), cfg)
    shards = sharding.shard_videos([t for t, _ in spec], world)
    # batch composition as data/dataset.py packed_batches builds it: longest video first (the host-space call
    # pipelines the batch by video group, and a group's recurrence lasts as long as its longest video)
    mine = sorted(shards[rank], key=lambda i: -spec[i][0])
    vids = [synth.make_video(spec[i][0], 1024, 128, spec[i][1]) for i in mine]
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    R = int(sum(lens))
    frames_global = sum(t for t, _ in spec)
    n_videos_global = len(spec)

    sd = synth.seeded_state_dict()
    model = AVBiLSTMModel(1024, 128, 512, attn_axis=axis).eval()
    model.load_state_dict(sd)
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
    visual_d, audio_d, pos_d = visual_h.to(dev), audio_h.to(dev), pos_h.to(dev)
    cps_list = [v.cps for v in vids]
    shots = ShotDesc([v.n_frames for v in vids], cps_list)     # packed once with the batch, like row_start / lengths
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # result gather (the only collective of the inference path): every rank knows every shard's shot count
    # from the host-side change points, so one padded all_gather of the keyshot picks per step suffices
    shots_per_rank = [sum(n_shots(*spec[i]) for i in sh) for sh in shards]
    pad = torch.zeros(max(shots_per_rank), dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * pad.numel(), dtype=torch.uint8, device=dev)

    def step_device():
        scores = nat.forward_rows(visual_d, audio_d, starts, lens, axis, "tf32")
        picks, seg_mean, summary, cps_start, _ = nat.summarize_rows(scores, pos_d, starts, lens, None, shots, 0.15)
        if world > 1:
            pad[:picks.numel()].copy_(picks)
            dist.all_gather_into_tensor(gathered, pad)
        return scores, picks, cps_start

    def step_host():
        # the public "score + summarise" call with pinned HOST buffers: features cross PCIe inside the call
        # (pipelined by video group), scores / picks / shot means / keyshot bitmap come back to host memory
        scores, picks, seg_mean, summary, _, _ = nat.score_and_summarize_rows(
            visual_h, audio_h, pos_h, starts, lens, None, shots, 0.15, axis, "tf32")
        return scores, picks, seg_mean, summary

    def stream_host(k):
        # the same call as a dataset loop makes it (scripts/evaluate.py:12-18 over packed batches): k batches through
        # evaluation.summary.summarize_stream, two in flight, so batch i+1 crosses PCIe while batch i is computed;
        # every batch's features go host -> device and its scores / picks / shot means / bitmap come back
        last = None
        for last in summarize_stream(model, ((visual_h, audio_h, pos_h, starts, lens, shots) for _ in range(k)),
                                     0.15, axis):
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
        out_dev = step_device()
        b.record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = _cabi.launch_count() - launches0
    stage_ms = _cabi.profile_read()
    _cabi.profile(0)
    ms_local = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    ms_per_step = max_over_ranks([ms_local], dev, world, dist)[0]
    # The timed step runs the pipelined tail (every group's value projection / out_proj / score head behind its own
    # recurrence, on the SMs the shorter groups have left): its stage timers see the recurrence stage (front done ->
    # last group's recurrence done) and the tail that remains behind the longest recurrence, not the individual tail
    # GEMMs, which overlap the recurrences.  Those are timed in a second, untimed-for-the-metric pass over the same
    # batch in the one-launch schedule (AVS_PIPE_TAIL=0: all recurrences, then each GEMM over all rows).
    stage_seq, seq_ms = None, None
    if stage_ms.get("tail_behind_longest_recurrence", (0.0, 0))[1] > 0:
        os.environ["AVS_PIPE_TAIL"] = "0"
        try:
            for _ in range(3):
                step_device()
            torch.cuda.synchronize()
            _cabi.profile(2)
            ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for a, b in ev2:
                flush.fill_(1)
                a.record()
                step_device()
                b.record()
            torch.cuda.synchronize()
            stage_seq = _cabi.profile_read()
            _cabi.profile(0)
            seq_ms = sum(a.elapsed_time(b) for a, b in ev2) / args.steps
        finally:
            os.environ.pop("AVS_PIPE_TAIL", None)
    scores_timed = out_dev[0].cpu()
    picks_timed, cps_start_timed = out_dev[1].cpu().numpy(), out_dev[2]

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
    # (2) the streamed form: K batches back to back, two in flight.  Every step moves its features through one of
    # two alternating device staging areas plus its activations -- far more than the 126 MB L2 -- so no explicit
    # flush is interleaved.
    # (the bare-copy floor is probed right before and right after the streamed steps: the host's memory system is
    # shared with other tenants and its delivered bandwidth drifts by +-15 % between moments on some boxes)
    floor_before = h2d_floor([visual_h, audio_h, pos_h], dev, world, dist)
    floor_modes_before = dict(h2d_floor.last)
    stream_host(4)
    barrier()
    e0 = time.perf_counter()
    res = stream_host(args.steps)
    barrier()
    e2e_ms_local = (time.perf_counter() - e0) / args.steps * 1e3
    e2e_ms, call_ms = max_over_ranks([e2e_ms_local, call_ms_local], dev, world, dist)
    scores_h, picks_h, segm_h, summ_h = res
    # (3) the same stream with the opt-in 16-bit host feature cache (data.dataset.packed_batches(feature_dtype="fp16"),
    # avs_model_set_feature_format): half the PCIe bytes per frame; fp32 features stay the primary number
    visual_h16, audio_h16 = visual_h.to(torch.float16).pin_memory(), audio_h.to(torch.float16).pin_memory()

    def stream_host16(k):
        last = None
        for last in summarize_stream(model, ((visual_h16, audio_h16, pos_h, starts, lens, shots) for _ in range(k)),
                                     0.15, axis):
            pass
        return last[:4]

    stream_host16(4)
    barrier()
    e0 = time.perf_counter()
    res16 = stream_host16(args.steps)
    barrier()
    e2e16_ms = max_over_ranks([(time.perf_counter() - e0) / args.steps * 1e3], dev, world, dist)[0]
    scores_h16 = res16[0].clone()
    h2d = R * (1024 + 128) * 4 + R * 4           # features + frame positions
    d2h = R * 4 + picks_h.numel() + segm_h.numel() * 8 + summ_h.numel()
    floor_after = h2d_floor([visual_h, audio_h, pos_h], dev, world, dist)
    floor_modes = {k: min(v, floor_modes_before.get(k, v)) for k, v in h2d_floor.last.items()}
    floor_ms = min(floor_before, floor_after)
    h2d16 = R * (1024 + 128) * 2 + R * 4
    floor16_ms = h2d_floor([visual_h16, audio_h16, pos_h], dev, world, dist)
    # clocks / throttle reasons sampled over ALL timed regions (device-resident steps, single calls, streamed steps)
    clocks = sampler.stop() if sampler else None

    peaks = load_peaks()
    attention = None
    if cfg == "infer" and not args.no_attention_probe:
        del flush
        torch.cuda.empty_cache()
        attention = attention_probe(sd, dev, peaks, world, dist)

    if rank != 0:
        finish(world, dist)
        return

    # ---- roofline of the dominant kernel (CUDA events inside the library, same timed region)
    mean_T = float(np.mean([t * t for t in lens]) / np.mean(lens))  # frame-weighted mean length
    flops = stage_flops_per_frame(axis, mean_T)
    kernels = kernel_table(stage_ms, args.steps, flops, R, peaks, 8 * R + 16 * sum(len(c) for c in cps_list))
    if stage_seq is not None:
        seq_tab = kernel_table(stage_seq, args.steps, flops, R, peaks, 8 * R + 16 * sum(len(c) for c in cps_list))
        for name, rec in seq_tab.items():
            if name not in kernels:
                rec["schedule"] = "one-launch pass (AVS_PIPE_TAIL=0), overlapped with the recurrences in the timed step"
                kernels[name] = rec
        kernels["lstm_recurrence"]["one_launch_schedule_ms_per_step"] = seq_tab["lstm_recurrence"]["ms_per_step"]
    roofline = roofline_of(kernels, ms_per_step, peaks, world)
    # the whole step against both rooflines: algorithmic FLOP of the layers the step executes and the
    # algorithmic bytes of its inputs / outputs (features read once, one score per frame, weights once per step)
    step_flop = R * sum(flops[k] for k in ("fc_gemm", "lstm_input_gemm", "lstm_recurrence", "attn_in_proj_gemm",
                                           "attention_core", "attn_out_proj_gemm", "score_head_gemm"))
    step_bytes = R * ((1024 + 128) * 4 + 4) + 32_035_332
    whole_step = {"algorithmic_tflop": step_flop / 1e12, "tflops": step_flop / (ms_per_step * 1e-3) / 1e12,
                  "frac_of_tensor_peak": step_flop / (ms_per_step * 1e-3) / 1e12 / peaks["tflops"],
                  "ideal_ms_at_tensor_peak": step_flop / (peaks["tflops"] * 1e12) * 1e3,
                  "algorithmic_bytes": step_bytes, "ideal_ms_at_hbm_peak": step_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3,
                  "note": "the step is bound by neither roofline: the recurrence is a chain of max(T) dependent "
                          "steps (roofline.share_of_step of the time at a few percent of the tensor peak); the "
                          "contractions around it run at kernels{}.frac_of_tensor_peak"}

    line = {
        "metric": METRIC, "value": frames_global / (ms_per_step * 1e-3), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg], "videos_per_gpu": len(vids), "frames_per_gpu": R,
                   "global_videos": n_videos_global, "attn_axis": axis, "l2": "256 MiB flush between timed steps",
                   "batch_order": "longest video first (packed_batches)", "parallelism": f"dp{world} by video"},
        "videos_per_s": n_videos_global / (ms_per_step * 1e-3),
        "e2e": {"value": frames_global / (e2e_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "h2d_floor_ms": floor_ms, "h2d_floor_before_after_ms": [floor_before, floor_after],
                "h2d_floor_by_piece_ms": floor_modes,
                "h2d_floor_note": "bare cudaMemcpyAsync of the step's own pinned input buffers (h2d_bytes_per_step), all ranks at once, "
                                  "whole buffers and 16 MB / 4 MB pieces, on the current and on a side stream (best of all), "
                                  "max over ranks: the end-to-end step cannot be shorter on this box",
                "frac_of_h2d_floor": floor_ms / e2e_ms,
                "mode": "streamed: evaluation.summary.summarize_stream, two batches in flight "
                        "(avs_forward_summarize_async); every step's H2D and D2H inside the timed region",
                "single_call_ms": call_ms, "single_call_value": frames_global / (call_ms * 1e-3),
                "l2": "per step the features go through alternating staging slots + activations >> L2; "
                      "single_call: 256 MiB flush before every call"},
        "e2e_fp16_features": {"value": frames_global / (e2e16_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e16_ms,
                              "h2d_bytes_per_step": int(h2d16), "d2h_bytes_per_step": int(d2h),
                              "h2d_floor_ms": floor16_ms, "frac_of_h2d_floor": floor16_ms / e2e16_ms,
                              "mode": "the streamed e2e step with the opt-in 16-bit host feature cache "
                                      "(packed_batches(feature_dtype='fp16') + avs_model_set_feature_format); "
                                      "secondary number, fp32 features (e2e) stay primary"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
        "whole_step": whole_step,
        "wall_s_timed_region": wall,
        "schedule": ({"timed_step": "pipelined tail: per recurrence group, recurrence -> attention (value projection, or "
                                    "q|k|v projection + attention core over the group's videos) -> out_proj -> score head "
                                    "on the group's own stream (scores bit-identical to the one-launch schedule)",
                      "one_launch_schedule_ms_per_step": seq_ms} if seq_ms is not None else
                     {"timed_step": "one launch per stage over all rows"}),
        "comm": {"backend": "nccl" if world > 1 else None, "nranks": world,
                 "collective": "all_gather_into_tensor of the keyshot picks, once per step" if world > 1 else None},
    }
    if attention is not None:
        line["attention"] = attention

    # ---- CPU baseline (reference CPU path port) on this box's host cores, N=1 only -- and parity of the timed
    # step's outputs against it
    if world == 1 and not args.no_cpu_baseline:
        from oracle import av_oracle, av_oracle_torch
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
        port.load_state_dict(sd)
        base_vids = build_global_batch(1, cfg)
        cpu_axis = "temporal" if axis == "temporal" else "literal"
        cpu_out = {}

        def cpu_step():
            sc = av_oracle_torch.run_videos(port, [(v.visual, v.audio) for v in base_vids], cpu_axis)
            for i, (v, s) in enumerate(zip(base_vids, sc)):
                cpu_out[i] = (s.numpy(), av_oracle.generate_summary(s.numpy(), v.cps, v.n_frames, v.positions)[0])

        cpu_step()
        reps, c0 = 0, time.perf_counter()
        while reps < 2 or (time.perf_counter() - c0 < 10.0 and reps < 20):
            cpu_step()
            reps += 1
        cdt = (time.perf_counter() - c0) / reps
        frames_cpu = sum(v.T for v in base_vids)
        line["cpu_baseline"] = {"value": frames_cpu / cdt, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"the full {len(base_vids)}-video batch x {reps} repetitions, B=1 loop "
                                          "(scripts/evaluate.py:12-18) over torch CPU ops + numpy summary oracle"}
        # parity: rank 0's videos are base_vids[mine[k]] (N = 1: the shard is the whole batch)
        worst, same = 0.0, 0
        for k, i in enumerate(mine):
            want_s, want_p = cpu_out[i]
            got_s = scores_timed[starts[k]:starts[k] + lens[k]].numpy()
            worst = max(worst, float(np.max(np.abs(got_s.astype(np.float64) - want_s) / np.abs(want_s))))
            same += int(np.array_equal(picks_timed[cps_start_timed[k]:cps_start_timed[k + 1]], want_p))
        line["parity"] = {"parity_max_rel_err": worst, "tolerance": 1e-3, "videos": len(mine),
                          "keyshot_selections_identical": same,
                          "against": "oracle/av_oracle_torch.py (bit-identical torch CPU port of the reference) + "
                                     "oracle/av_oracle.py summary, same inputs and weights as the timed step"}
        line["parity_max_rel_err"] = worst
        worst16 = 0.0
        for k, i in enumerate(mine):
            got_s = scores_h16[starts[k]:starts[k] + lens[k]].numpy()
            worst16 = max(worst16, float(np.max(np.abs(got_s.astype(np.float64) - cpu_out[i][0]) / np.abs(cpu_out[i][0]))))
        line["e2e_fp16_features"]["parity_max_rel_err"] = worst16
    print(json.dumps(line), flush=True)
    finish(world, dist)


# ------------------------------------------------------------------------------------ our arm: training config
def run_train(args, rank, world, local_rank):
    from avsum_b200 import _cabi, training
    from avsum_b200.models.av_model import AVBiLSTMModel
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = setup_dist(world, dev)
    B, T = 8, 320
    g = torch.Generator().manual_seed(100 + rank)
    visual_h = torch.randn(B, T, 1024, generator=g).pin_memory()
    audio_h = torch.randn(B, T, 128, generator=g).pin_memory()
    target_h = torch.rand(B, T, generator=g).pin_memory()
    visual, audio, target = visual_h.to(dev), audio_h.to(dev), target_h.to(dev)
    model = AVBiLSTMModel(1024, 128, 512, attn_axis="literal_b1")
    model.load_state_dict(synth.seeded_state_dict())
    model = model.to(dev).train()
    # the reference's optimiser (train_av_model.py:68: AdamW, lr 1e-4) in torch's single-kernel form: the default
    # (foreach) capturable AdamW launches ~100 tiny kernels per step for the per-parameter bias corrections
    # (profiles/r02i_train_launches.csv: 83 + 18 launches, 0.6 ms serialised); same update rule
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, capturable=True, **({} if args.foreach_adamw else {"fused": True}))
    n_params = sum(p.numel() for p in model.parameters())
    stepper = training.TrainStep(model, opt, torch.nn.functional.mse_loss, visual, audio, target,
                                 world_size=world, graph=not args.no_graph, warmup=max(args.warmup, 3),
                                 **({} if args.bucket_mb is None else {"bucket_mb": args.bucket_mb}))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        stepper(visual, audio, target)
    barrier()
    launches0 = _cabi.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = stepper(visual, audio, target)
    e1.record()
    barrier()
    # a replayed CUDA graph launches the kernels captured once: count those per replay
    launches = (_cabi.launch_count() - launches0) + (stepper.launches_per_replay * args.steps if stepper.graphed else 0)
    ms = max_over_ranks([e0.elapsed_time(e1) / args.steps], dev, world, dist)[0]
    # end to end: the batch comes from pinned host memory every step, the loss goes back to the host
    for _ in range(2):
        float(stepper(visual_h, audio_h, target_h).detach())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = float(stepper(visual_h, audio_h, target_h).detach())
    barrier()
    e2e_ms = max_over_ranks([(time.perf_counter() - t0) / args.steps * 1e3], dev, world, dist)[0]
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        peaks = load_peaks()
        flops = 3 * 15_990_912 * B * T    # forward + ~2x backward (SURVEY 8d per-frame figure, literal mode), per GPU
        tf = flops / (ms * 1e-3) / 1e12
        line = {"metric": "frames/sec trained", "value": B * T * world / (ms * 1e-3), "unit": "frames/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                "config": {"workload": WORKLOADS["train"], "videos_per_gpu": B, "frames_per_gpu": B * T,
                           "parameters": n_params, "gradient_bucket_mb": n_params * 4 / 1e6,
                           "cuda_graph": stepper.graphed, "allreduce": stepper.allreduce_mode,
                           "optimizer": "torch.optim.AdamW(lr=1e-4, capturable=True, "
                                        + ("foreach)" if args.foreach_adamw else "fused=True)"),
                           "parallelism": f"dp{world} by video", "l2": "activations + weights + optimiser state of a "
                           "step (~0.5 GB touched) exceed the 126 MB L2"},
                "steps_per_s": 1e3 / ms,
                "e2e": {"value": B * T * world / (e2e_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(B * T * (1024 + 128 + 1) * 4), "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"kernel": "whole step (~100 small kernels; BPTT chain is latency bound)", "bound": "tensor",
                             "achieved": tf, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": tf / peaks["tflops"],
                             "traffic": None},
                "final_loss": float(loss.detach()), "last_e2e_loss": last,
                "comm": {"backend": "nccl" if world > 1 else None, "nranks": world}}
        print(json.dumps(line), flush=True)
    finish(world, dist)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="infer", choices=["infer", "long", "train"])
    ap.add_argument("--axis", default="literal_b1", choices=["literal_b1", "temporal"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-attention-probe", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="--config train: launch the step eagerly")
    ap.add_argument("--foreach-adamw", action="store_true", help="--config train: torch's default (foreach) AdamW")
    ap.add_argument("--bucket-mb", type=float, default=None,
                    help="--config train: gradient all-reduce bucket size in MB (default: training.TrainStep's)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    if args.config == "train":
        run_train(args, rank, world, local_rank)
    else:
        run_infer(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
