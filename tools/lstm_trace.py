"""Phase trace of the tensor-core LSTM recurrence (debugging aid): where does one time step go?

    AVS_LSTM_TRACE=1 python tools/lstm_trace.py
"""
import ctypes as C
import os
os.environ.setdefault("AVS_PIPE_TAIL", "0")   # per-stage times / single launches: the one-launch schedule
import sys

os.environ["AVS_LSTM_TRACE"] = "1"
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa: E402,F401
from avsum_b200 import _cabi, synth  # noqa: E402
from avsum_b200.models.av_model import AVBiLSTMModel  # noqa: E402

NAMES = ["h landed -> fence.proxy.async + 16 MMAs + commit issued", "commit issued -> epilogue awake (MMA latency)",
         "tcgen05.ld (two 16-lane loads: all four gates of one unit per thread)",
         "+ input projection, cell update on the SFU, h staged as fp16", "tcgen05 fence + __syncwarp",
         "ld.shared of the staged chunk + st.async issue (one 16-byte message per lane)",
         "messages issued -> next h landed (DSMEM exchange, slowest peer)"]


def main():
    vids = synth.config2()
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    model = AVBiLSTMModel(1024, 128, 512, attn_axis="literal_b1").eval()
    model.load_state_dict(synth.seeded_state_dict())
    model = model.cuda()
    nat = model.native()
    visual = torch.cat([v.visual for v in vids]).cuda()
    audio = torch.cat([v.audio for v in vids]).cuda()
    for _ in range(3):
        nat.forward_rows(visual, audio, starts, lens, "literal_b1", "tf32")
    torch.cuda.synchronize()
    out = np.zeros(8, dtype=np.uint64)
    _cabi.check(_cabi.lib().avs_debug_lstm_trace(C.c_void_p(out.ctypes.data)))
    steps = int(out[7])
    tot = 0.0
    for n, v in zip(NAMES, out[:7]):
        per = float(v) / max(steps, 1)
        tot += per
        print(f"{per:8.1f} clk/step  {n}")
    print(f"{tot:8.1f} clk/step  total  ({steps} steps)")


if __name__ == "__main__":
    main()
