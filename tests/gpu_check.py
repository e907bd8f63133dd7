"""First-contact GPU checks, one subprocess per stage so a trap/hang in one kernel does not
hide the results of the others.  Run on the GPU box:

    python tests/gpu_check.py            # all stages, each under its own timeout
    python tests/gpu_check.py --stage gemm_tc

Prints one line per check; exits non-zero if any stage failed.
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["gemm_simt", "gemm_tc", "lstm", "forward_simt", "forward_tf32", "summarize", "f1", "config2"]


def rel_err(got, want):
    import torch
    return float((got.double() - want.double()).abs().max() / want.double().abs().max().clamp_min(1e-30))


def stage_gemm(prec):
    import torch
    from avsum_b200 import runtime
    torch.manual_seed(0)
    shapes = [(128, 128, 32), (128, 128, 128), (300, 512, 128), (1000, 2048, 512), (500, 64, 1024), (320, 512, 296),
              (21477, 3072, 1024), (77, 1024, 1024)]
    ok = True
    for (M, N, K) in shapes:
        x = torch.randn(M, K, device="cuda")
        w = torch.randn(N, K, device="cuda") / K ** 0.5
        b = torch.randn(N, device="cuda")
        for relu in (False, True):
            want = torch.nn.functional.linear(x.double(), w.double(), b.double())
            if relu:
                want = want.relu()
            got = runtime.linear(x, w, b, relu=relu, precision=prec)
            torch.cuda.synchronize()
            e = rel_err(got, want)
            tol = 2e-6 if prec == "fp32_simt" else 2e-3
            flag = "ok" if e < tol else "FAIL"
            ok &= e < tol
            print(f"  linear[{prec}] M={M} N={N} K={K} relu={int(relu)} rel_err={e:.3e} {flag}")
    return ok


def stage_lstm():
    import numpy as np
    import torch
    from avsum_b200 import runtime, synth, _cabi
    import ctypes as C
    sd = synth.seeded_state_dict()
    nat = runtime.NativeModel({k: v.cuda() for k, v in sd.items()}, 1024, 128)
    torch.manual_seed(1)
    ok = True
    for lens in ([37], [5, 64, 1, 33], list(range(20, 39))):
        R = sum(lens)
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
        v = torch.randn(R, 512)
        a = torch.randn(R, 512)
        lv = torch.nn.LSTM(512, 256, bidirectional=True, batch_first=True)
        la = torch.nn.LSTM(512, 256, bidirectional=True, batch_first=True)
        lv.load_state_dict({k.split(".", 1)[1]: t for k, t in sd.items() if k.startswith("visual_bilstm")})
        la.load_state_dict({k.split(".", 1)[1]: t for k, t in sd.items() if k.startswith("audio_bilstm")})
        want = torch.empty(R, 1024)
        with torch.no_grad():
            for s, n in zip(starts, lens):
                want[s:s + n, :512] = lv(v[None, s:s + n])[0][0]
                want[s:s + n, 512:] = la(a[None, s:s + n])[0][0]
        for prec in ("fp32_simt", "tf32"):
            fused = torch.zeros(R, 1024, device="cuda")
            vc, ac = v.cuda(), a.cuda()
            ln = np.asarray(lens, dtype=np.int32)
            _cabi.check(nat.lib.avs_bilstm_pair(nat._handle, C.c_void_p(vc.data_ptr()), C.c_void_p(ac.data_ptr()), R,
                                                len(lens), _cabi.np_ptr(starts), _cabi.np_ptr(ln),
                                                _cabi.PRECISIONS[prec], C.c_void_p(fused.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
            e = float((fused.cpu() - want).abs().max())
            tol = 2e-5 if prec == "fp32_simt" else 3e-3
            ok &= e < tol
            print(f"  bilstm[{prec}] lens={lens if len(lens) < 6 else str(lens[:3]) + '...'} max_abs_err={e:.3e} {'ok' if e < tol else 'FAIL'}")
    return ok


def stage_forward(prec):
    import numpy as np
    import torch
    from avsum_b200 import synth
    from avsum_b200.models.av_model import AVBiLSTMModel
    ok = True
    tol = 2e-5 if prec == "fp32_simt" else 1e-3
    gdir = os.path.join(ROOT, "tests", "golden")
    for spread in (0, 1):
        g = np.load(os.path.join(gdir, f"config1_spread{spread}.npz"))
        sd = synth.seeded_state_dict(spread=bool(spread))
        assert abs(synth.state_dict_checksum(sd) - float(g["weights_checksum"])) < 1e-6
        m = AVBiLSTMModel(1024, 128, 512, precision=prec).eval()
        m.load_state_dict(sd)
        m = m.cuda()
        vid = synth.config1()
        for axis in ("literal", "temporal"):
            got = m(vid.visual[None].cuda(), vid.audio[None].cuda(), attn_axis=axis).cpu().numpy()
            want = g["scores_" + axis]
            e = float(np.max(np.abs(got - want) / np.abs(want)))
            ok &= e < tol
            print(f"  forward[{prec}] config1 spread={spread} {axis:8s} max_rel_err={e:.3e} {'ok' if e < tol else 'FAIL'}")
    for name in ("batch3_T17", "batch2_T1", "batch1_T1", "default_dims_T40", "batch2_T130_spread"):
        g = np.load(os.path.join(gdir, name + ".npz"))
        vd, ad, B, T = int(g["visual_dim"]), int(g["audio_dim"]), int(g["B"]), int(g["T"])
        sd = synth.seeded_state_dict(vd, ad, 512, 0, bool(int(g["spread"])))
        m = AVBiLSTMModel(vd, ad, 512, precision=prec).eval()
        m.load_state_dict(sd)
        m = m.cuda()
        gen = torch.Generator().manual_seed(int(g["seed_in"]))
        visual = torch.randn(B, T, vd, generator=gen)
        audio = torch.randn(B, T, ad, generator=gen)
        for axis in ("literal", "temporal"):
            got = m(visual.cuda(), audio.cuda(), attn_axis=axis).cpu().numpy()
            want = g["scores_" + axis]
            shape_ok = got.shape == want.shape
            e = float(np.max(np.abs(got - want) / np.abs(want))) if shape_ok else float("inf")
            ok &= shape_ok and e < tol
            print(f"  forward[{prec}] {name:20s} {axis:8s} shape={got.shape} max_rel_err={e:.3e} {'ok' if e < tol and shape_ok else 'FAIL'}")
    return ok


def stage_summarize():
    import numpy as np
    import torch
    from avsum_b200 import synth, runtime
    from oracle import av_oracle
    sd = synth.seeded_state_dict()
    nat = runtime.NativeModel({k: v.cuda() for k, v in sd.items()}, 1024, 128)
    vids = synth.config2()
    rng = np.random.default_rng(7)
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    ok = True
    for trial, kind in enumerate(["uniform", "near_tie", "spiky"]):
        if kind == "uniform":
            scores = rng.random(sum(lens)).astype(np.float32)
        elif kind == "near_tie":
            scores = (0.52 + 1e-4 * rng.standard_normal(sum(lens))).astype(np.float32)
        else:
            scores = (rng.random(sum(lens)) ** 8).astype(np.float32)
        pos = np.concatenate([v.positions for v in vids]).astype(np.int32)
        for space in ("cuda", "cpu"):
            s_t = torch.from_numpy(scores).to(space)
            p_t = torch.from_numpy(pos).to(space)
            picks, seg_mean, summary, cps_start, sum_start = nat.summarize_rows(
                s_t, p_t, starts, lens, [v.n_frames for v in vids], [v.cps for v in vids], 0.15)
            torch.cuda.synchronize()
            picks, seg_mean, summary = picks.cpu().numpy(), seg_mean.cpu().numpy(), summary.cpu().numpy()
            bad = 0
            for i, v in enumerate(vids):
                wp, ws, wm = av_oracle.generate_summary(scores[starts[i]:starts[i] + lens[i]], v.cps, v.n_frames, v.positions)
                bad += int(not np.array_equal(wp, picks[cps_start[i]:cps_start[i + 1]]))
                bad += int(not np.array_equal(wm, seg_mean[cps_start[i]:cps_start[i + 1]]))
                bad += int(not np.array_equal(ws, summary[sum_start[i]:sum_start[i + 1]]))
            ok &= bad == 0
            print(f"  summarize {kind:9s} space={space:4s} mismatching arrays={bad} picked={int(picks.sum())}/{picks.size} {'ok' if bad == 0 else 'FAIL'}")
    return ok


def stage_f1():
    import numpy as np
    from avsum_b200.evaluation.metrics import compute_temporal_f1_batch
    from oracle import av_oracle
    rng = np.random.default_rng(3)
    preds, gts = [], []
    for _ in range(40):
        def shots():
            e = np.sort(rng.choice(5000, size=2 * int(rng.integers(1, 30)), replace=False))
            return [(int(e[2 * i]), int(e[2 * i + 1])) for i in range(len(e) // 2)]
        preds.append(shots())
        gts.append(shots())
    got = compute_temporal_f1_batch(preds, gts)
    want = np.asarray([av_oracle.temporal_f1(p, g) for p, g in zip(preds, gts)])
    same = bool(np.array_equal(got, want))
    print(f"  temporal_f1 40 videos bit-exact={same} max_abs={np.max(np.abs(got - want)):.3e}")
    return same


def stage_config2():
    import numpy as np
    import torch
    from avsum_b200 import synth
    from avsum_b200.models.av_model import AVBiLSTMModel
    from avsum_b200.evaluation.summary import summarize_videos
    g = np.load(os.path.join(ROOT, "tests", "golden", "config2_first4_spread1.npz"))
    sd = synth.seeded_state_dict(spread=True)
    ok = True
    vids = synth.config2()
    for axis in ("literal_b1", "temporal"):
        m = AVBiLSTMModel(1024, 128, 512, attn_axis=axis).eval()
        m.load_state_dict(sd)
        m = m.cuda()
        t0 = time.time()
        res = summarize_videos(m, vids)
        torch.cuda.synchronize()
        t1 = time.time()
        res = summarize_videos(m, vids)
        torch.cuda.synchronize()
        t2 = time.time()
        got = np.concatenate([r.scores.numpy() for r in res[:4]])
        want = g["scores_literal" if axis == "literal_b1" else "scores_temporal"]
        e = float(np.max(np.abs(got - want) / np.abs(want)))
        ok &= e < 1e-3
        print(f"  config2 {axis:10s} first4 max_rel_err={e:.3e} wall first={t1 - t0:.3f}s second={t2 - t1:.3f}s "
              f"({sum(v.T for v in vids) / (t2 - t1):.0f} frames/s host-in host-out) {'ok' if e < 1e-3 else 'FAIL'}")
    return ok


def run_stage(name):
    import torch
    import avsum_b200  # noqa: F401
    from avsum_b200 import _cabi
    assert torch.cuda.is_available(), "no CUDA device"
    assert _cabi.lib().avs_device_ok() == 1, "avs_device_ok() == 0"
    if name == "gemm_simt":
        return stage_gemm("fp32_simt")
    if name == "gemm_tc":
        return stage_gemm("tf32")
    if name == "lstm":
        return stage_lstm()
    if name == "forward_simt":
        return stage_forward("fp32_simt")
    if name == "forward_tf32":
        return stage_forward("tf32")
    if name == "summarize":
        return stage_summarize()
    if name == "f1":
        return stage_f1()
    if name == "config2":
        return stage_config2()
    raise SystemExit(f"unknown stage {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default=None)
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.stage:
        ok = run_stage(args.stage)
        sys.exit(0 if ok else 1)
    failed = []
    for st in STAGES:
        print(f"== stage {st}", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], timeout=args.timeout)
            rc = r.returncode
        except subprocess.TimeoutExpired:
            rc = "timeout"
        print(f"== stage {st} rc={rc} ({time.time() - t0:.1f}s)", flush=True)
        if rc != 0:
            failed.append(st)
    print("FAILED STAGES:", failed if failed else "none")
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
