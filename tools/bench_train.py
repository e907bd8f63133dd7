"""BASELINE configs[4]: synthetic AVModel training step, batch-sharded data parallelism, NCCL gradient all-reduce.

    python tools/bench_train.py [--steps K] [--warmup W] [--videos 8] [--frames 320]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...

One step = scripts/train_av_model.py:86-96 for a batch of B = 1 samples per GPU: forward in train mode
(Dropout(0.3) active), F.mse_loss against rand targets, backward (native kernels, avsum_b200/training.py),
one flat-bucket gradient all-reduce (N > 1), torch.optim.AdamW(lr=1e-4).  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa: E402,F401
from avsum_b200 import synth, training  # noqa: E402
from avsum_b200.models.av_model import AVBiLSTMModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--videos", type=int, default=8)
    ap.add_argument("--frames", type=int, default=320)
    ap.add_argument("--cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the step from a CUDA graph (training.GraphedTrainStep)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        # (also at WARN); NCCL logging is therefore off unless AVS_NCCL_DEBUG asks for it, and then goes to a file
        os.environ.pop("NCCL_DEBUG", None)
        if os.environ.get("AVS_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = os.environ["AVS_NCCL_DEBUG"]
            os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/avs_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
    B, T = args.videos, args.frames
    g = torch.Generator().manual_seed(100 + rank)
    visual = torch.randn(B, T, 1024, generator=g).to(dev)
    audio = torch.randn(B, T, 128, generator=g).to(dev)
    target = torch.rand(B, T, generator=g).to(dev)
    model = AVBiLSTMModel(1024, 128, 512, attn_axis="literal_b1")
    model.load_state_dict(synth.seeded_state_dict())
    model = model.to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, capturable=args.graph)
    n_params = sum(p.numel() for p in model.parameters())
    graphed = None
    if args.graph and world > 1:
        raise SystemExit("--graph is single-GPU only: the captured step does not include the NCCL gradient all-reduce")
    if args.graph:
        graphed = training.GraphedTrainStep(model, opt, torch.nn.functional.mse_loss, visual, audio, target,
                                            allreduce=world > 1, warmup=max(args.warmup, 3))

    def step():
        if graphed is not None:
            return graphed(visual, audio, target)
        preds = model(visual, audio)
        loss = torch.nn.functional.mse_loss(preds, target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            training.allreduce_gradients(model.parameters())
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        flops = 3 * (15_990_912) * B * T * world      # forward + ~2x backward (SURVEY 8d per-frame figure, literal mode)
        line = {"config": f"config5: {B} videos/GPU x T={T}, train mode (dropout 0.3), mse_loss, AdamW lr 1e-4, "
                          f"{n_params} parameters, flat fp32 gradient all-reduce ({n_params * 4 / 1e6:.1f} MB)",
                "n_gpus": world, "cuda_graph": bool(args.graph), "ms_per_step": ms, "steps_per_s": 1e3 / ms, "frames_per_s": B * T * world / (ms * 1e-3),
                "approx_tflops": flops / (ms * 1e-3) / 1e12, "final_loss": float(loss)}
        if args.cpu_baseline and world == 1:
            import torch.nn as nn

            class RefModel(nn.Module):   # the reference's module tree and forward (models/av_model.py:7-46), CPU timing only
                def __init__(self, vd, ad, hd):
                    super().__init__()
                    self.visual_fc = nn.Sequential(nn.Linear(vd, hd), nn.ReLU(), nn.Dropout(0.3))
                    self.audio_fc = nn.Sequential(nn.Linear(ad, hd), nn.ReLU(), nn.Dropout(0.3))
                    self.visual_bilstm = nn.LSTM(hd, hd // 2, bidirectional=True, batch_first=True)
                    self.audio_bilstm = nn.LSTM(hd, hd // 2, bidirectional=True, batch_first=True)
                    self.attention = nn.MultiheadAttention(embed_dim=hd * 2, num_heads=4)
                    self.scorer = nn.Sequential(nn.Linear(hd * 2, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())

                def forward(self, visual, audio):
                    v, _ = self.visual_bilstm(self.visual_fc(visual))
                    a, _ = self.audio_bilstm(self.audio_fc(audio))
                    fused = torch.cat([v, a], dim=-1)
                    return self.scorer(self.attention(fused, fused, fused)[0]).squeeze()

            torch.set_num_threads(os.cpu_count() or 1)
            port = RefModel(1024, 128, 512).train()
            port.load_state_dict(synth.seeded_state_dict())
            popt = torch.optim.AdamW(port.parameters(), lr=1e-4)
            vh, ah, th = visual.cpu(), audio.cpu(), target.cpu()

            def cpu_step():
                popt.zero_grad()
                for b in range(B):      # the reference's loop: one video per forward/backward
                    l = torch.nn.functional.mse_loss(port(vh[b:b + 1], ah[b:b + 1]), th[b]) / B
                    l.backward()
                popt.step()

            cpu_step()
            t0 = time.perf_counter()
            for _ in range(2):
                cpu_step()
            cdt = (time.perf_counter() - t0) / 2
            line["cpu_reference_ms_per_step"] = cdt * 1e3
            line["cpu_cores"] = os.cpu_count()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
