"""A/B of two builds of libavsum_b200.so on the same box: runs tools/lstm_scaling.py-style timings (per-stage CUDA
events inside the library) once per library, alternating, in fresh processes.

    python tools/ab_lib.py build/libavsum_b200_old.so [more.so ...] [n_videos ...]
"""
import os, subprocess, sys
os.environ.setdefault("AVS_PIPE_TAIL", "0")   # per-stage times / single launches: the one-launch schedule
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
import numpy as np, torch
sys.path.insert(0, %r)
import avsum_b200
from avsum_b200 import synth, _cabi
from avsum_b200.models.av_model import AVBiLSTMModel
vids = sorted(synth.config2(), key=lambda v: -v.T)
model = AVBiLSTMModel(1024, 128, 512, attn_axis="literal_b1").eval()
model.load_state_dict(synth.seeded_state_dict(spread=True))
model = model.cuda()
nat = model.native()
for n in [int(x) for x in sys.argv[1:]]:
    sub = vids[:n]
    lens = [v.T for v in sub]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    v = torch.cat([x.visual for x in sub]).cuda(); a = torch.cat([x.audio for x in sub]).cuda()
    for _ in range(3):
        out = nat.forward_rows(v, a, starts, lens, "literal_b1", "tf32")
    torch.cuda.synchronize()
    _cabi.profile(2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = nat.forward_rows(v, a, starts, lens, "literal_b1", "tf32")
    e1.record(); torch.cuda.synchronize()
    st = _cabi.profile_read(); _cabi.profile(0)
    ms = st["lstm_recurrence"][0] / 20
    print(f"  {n:3d} videos: lstm {ms:.4f} ms = {ms * 1e3 / max(lens):.4f} us/step, forward {e0.elapsed_time(e1) / 20:.4f} ms, "
          f"checksum {float(out.double().sum()):.9f}", flush=True)
    if n == 50:
        print("      stages (ms): " + ", ".join(f"{k} {v[0] / 20:.4f}" for k, v in st.items() if v[1]), flush=True)
        for _ in range(3):
            out = nat.forward_rows(v, a, starts, lens, "temporal", "tf32")
        torch.cuda.synchronize()
        _cabi.profile(2)
        for _ in range(20):
            out = nat.forward_rows(v, a, starts, lens, "temporal", "tf32")
        st = _cabi.profile_read(); _cabi.profile(0)
        print("      temporal stages (ms): " + ", ".join(f"{k} {v[0] / 20:.4f}" for k, v in st.items() if v[1])
              + f", checksum {float(out.double().sum()):.9f}", flush=True)
''' % ROOT

others = [os.path.abspath(a) for a in sys.argv[1:] if a.endswith(".so")]
ns = [a for a in sys.argv[1:] if not a.endswith(".so")] or ["1", "8", "50"]
for rep in range(2):
    for name, path in [(os.path.basename(o), o) for o in others] + [("in-tree", None)]:
        env = dict(os.environ)
        env.pop("AVS_LIB_PATH", None)
        if path:
            env["AVS_LIB_PATH"] = path
        print(f"[{name}] {path or 'in-tree build'}", flush=True)
        subprocess.run([sys.executable, "-c", CHILD] + ns, env=env, check=True)
