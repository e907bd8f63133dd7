"""LSTM recurrence time vs number of videos at (almost) fixed max length: is a step bound by its own dependency
chain or by contention between the chains that share an SM?"""
import os, sys
os.environ.setdefault("AVS_PIPE_TAIL", "0")   # per-stage times / single launches: the one-launch schedule
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa
from avsum_b200 import synth, _cabi
from avsum_b200.models.av_model import AVBiLSTMModel

vids = sorted(synth.config2(), key=lambda v: -v.T)
model = AVBiLSTMModel(1024, 128, 512, attn_axis="literal_b1").eval()
model.load_state_dict(synth.seeded_state_dict())
model = model.cuda()
nat = model.native()
for n in (1, 4, 8, 16, 24, 28, 32, 36, 40, 44, 48, 50):
    sub = vids[:n]
    lens = [v.T for v in sub]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    v = torch.cat([x.visual for x in sub]).cuda()
    a = torch.cat([x.audio for x in sub]).cuda()
    for _ in range(3):
        nat.forward_rows(v, a, starts, lens, "literal_b1", "tf32")
    torch.cuda.synchronize()
    _cabi.profile(2)
    for _ in range(5):
        nat.forward_rows(v, a, starts, lens, "literal_b1", "tf32")
    st = _cabi.profile_read()
    _cabi.profile(0)
    ms = st["lstm_recurrence"][0] / 5
    print(f"{n:3d} videos, max len {max(lens)}, min len {min(lens)}: lstm {ms:.3f} ms = {ms * 1e3 / max(lens):.3f} us/step")
