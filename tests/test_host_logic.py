"""CPU tests of the host-side mirror of the reference interface (no compute)."""
import os

import numpy as np
import pytest
import torch

from avsum_b200 import synth
from avsum_b200.models.av_model import AVBiLSTMModel, AVModel, AVSummarizer
from avsum_b200.models.attention import MultiHeadSelfAttention
from avsum_b200.utils.alignments import align_shots_to_annotations
from avsum_b200.utils.shot_metrics import calculate_overlap

# the 28 state_dict keys of the reference (measured on the imported class)
REFERENCE_KEYS = (
    ["visual_fc.0.weight", "visual_fc.0.bias", "audio_fc.0.weight", "audio_fc.0.bias"]
    + [f"{m}.{p}{s}" for m in ("visual_bilstm", "audio_bilstm") for s in ("", "_reverse")
       for p in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
    + ["attention.in_proj_weight", "attention.in_proj_bias", "attention.out_proj.weight", "attention.out_proj.bias",
       "scorer.0.weight", "scorer.0.bias", "scorer.2.weight", "scorer.2.bias"])


def test_state_dict_keys_and_shapes_match_reference():
    m = AVBiLSTMModel()
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(REFERENCE_KEYS) and len(sd) == 28
    assert sd["visual_fc.0.weight"].shape == (512, 4096) and sd["audio_fc.0.weight"].shape == (512, 296)
    assert sd["visual_bilstm.weight_ih_l0"].shape == (1024, 512) and sd["audio_bilstm.weight_hh_l0_reverse"].shape == (1024, 256)
    assert sd["attention.in_proj_weight"].shape == (3072, 1024) and sd["scorer.2.weight"].shape == (1, 64)
    assert sum(p.numel() for p in m.parameters()) == 9_667_713
    assert sum(p.numel() for p in AVBiLSTMModel(1024, 128, 512).parameters()) == 8_008_833
    assert AVModel is AVBiLSTMModel and AVSummarizer is AVBiLSTMModel


def test_seeded_construction_reproduces_reference_weights(golden_dir):
    g = np.load(os.path.join(golden_dir, "config1_spread0.npz"))
    torch.manual_seed(0)
    m = AVBiLSTMModel(1024, 128, 512)
    assert abs(synth.state_dict_checksum(m.state_dict()) - float(g["weights_checksum"])) < 1e-6


def test_mhsa_surface():
    a = MultiHeadSelfAttention(1024, 4)
    assert sorted(a.state_dict().keys()) == sorted(f"{n}.{p}" for n in ("query", "key", "value", "out") for p in ("weight", "bias"))
    assert a.num_heads == 4 and a.dim_head == 256


def test_training_mode_has_no_cpu_fallback_and_rejects_unsupported_attention():
    m = AVBiLSTMModel(1024, 128, 512).train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):       # the training step needs the CUDA kernels too
        m(torch.zeros(1, 2, 1024), torch.zeros(1, 2, 128))
    with pytest.raises(NotImplementedError):                         # cross-video attention has no backward
        m(torch.zeros(2, 2, 1024), torch.zeros(2, 2, 128))
    with pytest.raises(NotImplementedError):
        AVBiLSTMModel(1024, 128, 512, attn_axis="temporal").train()(torch.zeros(1, 2, 1024), torch.zeros(1, 2, 128))


def test_synthetic_workloads_are_deterministic():
    vids = synth.config2()
    assert len(vids) == 50 and sum(v.T for v in vids) == 21_477
    assert all(200 <= v.T <= 700 for v in vids)
    v0 = synth.config2()[0]
    assert torch.equal(v0.visual, vids[0].visual) and np.array_equal(v0.cps, vids[0].cps)
    for v in vids[:5]:
        assert v.cps[0, 0] == 0 and v.cps[-1, 1] == v.n_frames - 1
        assert np.all(v.cps[1:, 0] == v.cps[:-1, 1] + 1)
    c1 = synth.config1()
    assert c1.visual.shape == (320, 1024) and c1.audio.shape == (320, 128)


def test_alignment_and_overlap_helpers(golden_dir):
    g = np.load(os.path.join(golden_dir, "helpers.npz"))
    shots = [tuple(int(x) for x in s) for s in g["shots"]]
    got = align_shots_to_annotations(shots, g["ann"], 30.0)
    assert got.dtype == torch.float64 and np.array_equal(got.numpy(), g["aligned"])
    pred = [tuple(int(x) for x in p) for p in g["pred"]]
    gt = [tuple(int(x) for x in p) for p in g["gt"]]
    assert calculate_overlap(pred, gt) == int(g["overlap"])


def test_feature_dir_dataset_and_packed_batches(tmp_path):
    """data/dataset.py:8-32 contract: <dir>/<vid>/{visual,audio,scores}.npy -> (features dict, scores)."""
    import numpy as np
    import torch
    from avsum_b200.data.dataset import BaseDataset, packed_batches
    rng = np.random.default_rng(0)
    lens = {"vid_b": 7, "vid_a": 19, "vid_c": 3, "vid_d": 12}
    for name, t in lens.items():
        d = tmp_path / name
        d.mkdir()
        np.save(d / "visual.npy", rng.random((t, 16)).astype(np.float32))
        np.save(d / "audio.npy", rng.random((t, 4)).astype(np.float32))
        np.save(d / "scores.npy", rng.random(t).astype(np.float32))
    ds = BaseDataset(str(tmp_path))
    assert len(ds) == 4 and ds.video_ids == sorted(lens)
    feats, scores = ds[0]
    assert set(feats) == {"visual", "audio"} and feats["visual"].shape == (19, 16) and scores.shape == (19,)
    batches = list(packed_batches(ds, max_frames=30, pin=False))
    seen = sorted(i for b in batches for i in b.indices)
    assert seen == [0, 1, 2, 3]
    for b in batches:
        assert int(b.lengths.sum()) == b.visual.shape[0] == b.audio.shape[0] == b.scores.shape[0] <= 30
        assert list(b.lengths) == sorted(b.lengths, reverse=True)          # bucketed: longest first
        for k, i in enumerate(b.indices):
            s, n = int(b.row_start[k]), int(b.lengths[k])
            assert torch.equal(b.visual[s:s + n], ds[i][0]["visual"]) and torch.equal(b.scores[s:s + n], ds[i][1])
    # opt-in 16-bit feature cache: same batches, features packed as IEEE half (scores untouched)
    half = list(packed_batches(ds, max_frames=30, pin=False, feature_dtype="fp16"))
    assert [b.indices for b in half] == [b.indices for b in batches]
    for hb, fb in zip(half, batches):
        assert hb.visual.dtype == torch.float16 and hb.audio.dtype == torch.float16 and hb.scores.dtype == fb.scores.dtype
        assert torch.equal(hb.visual, fb.visual.to(torch.float16))
    with pytest.raises(ValueError):
        next(packed_batches(ds, feature_dtype="bf16"))


def test_shot_descriptors_pack_like_the_per_call_path():
    """runtime.ShotDesc (packed once by the loader) holds exactly what the per-call path derives from lists."""
    from avsum_b200.runtime import ShotDesc, _shots
    vids = synth.config2()[:7]
    sd = ShotDesc([v.n_frames for v in vids], [v.cps for v in vids])
    assert sd.cps.dtype == np.int32 and sd.cps.shape[1] == 2 and sd.cps.flags["C_CONTIGUOUS"]
    assert sd.cps_start[0] == 0 and list(np.diff(sd.cps_start)) == [len(v.cps) for v in vids]
    assert np.array_equal(sd.cps, np.concatenate([np.asarray(v.cps, np.int32).reshape(-1, 2) for v in vids]))
    assert list(np.diff(sd.summary_start)) == [v.n_frames for v in vids] and sd.summary_start.dtype == np.int64
    assert _shots(None, sd) is sd
    empty = ShotDesc([], [])
    assert empty.cps.shape == (0, 2) and list(empty.cps_start) == [0] and list(empty.summary_start) == [0]
    with pytest.raises(ValueError):
        ShotDesc([100, 200], [vids[0].cps])


def test_numa_binding_is_best_effort():
    """One process per GPU binds itself next to its GPU when sysfs exposes the topology; without a GPU (here), with
    one NUMA node or with AVS_NO_NUMA_BIND it must be a harmless no-op that says why."""
    import os
    from avsum_b200 import sharding
    before = os.sched_getaffinity(0)
    msg = sharding.bind_process_to_gpu_numa(0)
    assert msg.startswith("numa:")
    assert os.sched_getaffinity(0) == before or "bound to" in msg
    os.environ["AVS_NO_NUMA_BIND"] = "1"
    try:
        assert "disabled" in sharding.bind_process_to_gpu_numa(0)
    finally:
        del os.environ["AVS_NO_NUMA_BIND"]


def test_graphed_train_step_needs_cuda_and_a_capturable_optimiser():
    """training.GraphedTrainStep refuses host tensors (no CPU fallback) before it touches the model."""
    import torch
    from avsum_b200 import training
    lin = torch.nn.Linear(4, 1)
    opt = torch.optim.AdamW(lin.parameters(), lr=1e-3)
    with pytest.raises(RuntimeError, match="CUDA"):
        training.GraphedTrainStep(lin, opt, torch.nn.functional.mse_loss, torch.zeros(1, 2, 4), torch.zeros(1, 2, 4),
                                  torch.zeros(1, 2))


def test_bench_global_batch_spec_matches_the_synthetic_configs():
    """bench.py generates only a rank's shard of the global batch; the (length, seed) spec every rank derives must
    describe exactly the videos synth.config2 / config4 build (the CPU baseline and the parity check index by it)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("avs_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    one = bench.global_batch_spec(1, "infer")
    vids = synth.config2()
    assert [t for t, _ in one] == [v.T for v in vids] and sum(t for t, _ in one) == 21477
    v7 = synth.make_video(*one[7][:1], 1024, 128, one[7][1])
    assert np.array_equal(v7.cps, vids[7].cps) and bench.n_shots(*one[7]) == len(vids[7].cps)
    two = bench.global_batch_spec(2, "infer")
    assert two[:50] == one and len(two) == 100 and two[50][1] == 2234
    long = bench.global_batch_spec(2, "long")
    assert [t for t, _ in long] == [8192] * 16 and len({s for _, s in long}) == 16
    # stage FLOP table: the pipelined stage is the sum of its parts
    f = bench.stage_flops_per_frame("temporal", 400.0)
    assert f["frontend_lstm_pipelined"] == f["frontend_gemms"] + f["lstm_recurrence"]
    assert f["attention_core"] == 4 * 400.0 * 1024


# ---- row layout of the callers (longest video first) against a stand-in for the native handle ---------------------
# The callers below reorder the rows of a batch (so that every recurrence group owns one block of rows) while the
# per-video descriptors and results stay in the caller's order.  That bookkeeping is host logic; it is checked here
# with a stand-in for NativeModel whose "scores" are a row-local function of the features, so any mix-up between the
# row layout and the descriptors shows.  (The GPU tests check the same callers against the oracle.)

class _StubNative:
    def __init__(self):
        self.calls = []

    @staticmethod
    def _row_scores(visual, audio):
        return (visual[:, 0].double() * 0.25 + audio[:, 0].double() * 0.5 + 0.125).float()

    def forward_rows(self, visual, audio, row_start, lengths, attn_axis="literal", precision="tf32", out=None):
        rs, ln = np.asarray(row_start, np.int64), np.asarray(lengths, np.int64)
        # the descriptors must tile the rows exactly once, whatever the order of the videos
        cover = np.zeros(visual.shape[0], np.int32)
        for s, n in zip(rs, ln):
            cover[s:s + n] += 1
        assert np.all(cover == 1)
        self.calls.append((rs.copy(), ln.copy(), attn_axis))
        return self._row_scores(visual, audio)

    def score_and_summarize_rows(self, visual, audio, positions, row_start, lengths, n_frames, cps_list,
                                 proportion=0.15, attn_axis="literal_b1", precision="tf32"):
        scores = self.forward_rows(visual, audio, row_start, lengths, attn_axis, precision)
        rs, ln = np.asarray(row_start, np.int64), np.asarray(lengths, np.int64)
        cps_start = np.concatenate([[0], np.cumsum([len(c) for c in cps_list])]).astype(np.int64)
        sum_start = np.concatenate([[0], np.cumsum(n_frames)]).astype(np.int64)
        picks = torch.zeros(int(cps_start[-1]), dtype=torch.uint8)
        seg_mean = torch.zeros(int(cps_start[-1]), dtype=torch.int64)
        summary = torch.zeros(int(sum_start[-1]), dtype=torch.uint8)
        pos = positions.numpy()
        for k in range(len(ln)):
            tag = int(round(float(scores[rs[k]]) * 1e4))                 # a value only this video's rows produce
            assert pos[rs[k]] == 0 and pos[rs[k] + ln[k] - 1] == (ln[k] - 1) * synth.SAMPLE_STRIDE
            seg_mean[cps_start[k]:cps_start[k + 1]] = tag
            picks[cps_start[k]:cps_start[k + 1]] = tag % 2
            summary[sum_start[k]:sum_start[k + 1]] = tag % 251
        return scores, picks, seg_mean, summary, cps_start, sum_start


class _StubModel:
    attn_axis = "literal"
    precision = "tf32"

    def __init__(self):
        self.nat = _StubNative()

    def native(self):
        return self.nat

    def eval(self):
        return self

    def parameters(self):
        return iter([torch.zeros(1)])


def _mixed_videos(lengths, seed0=77):
    return [synth.make_video(t, 16, 4, seed0 + i) for i, t in enumerate(lengths)]


def test_summarize_videos_returns_results_in_the_callers_order():
    from avsum_b200.evaluation.summary import summarize_videos
    lengths = [9, 31, 9, 17, 40, 3, 17]                                  # unsorted, with ties
    vids = _mixed_videos(lengths)
    model = _StubModel()
    res = summarize_videos(model, vids)
    rs, ln, axis = model.nat.calls[0]
    assert axis == "literal_b1" and list(ln) == sorted(lengths, reverse=True)      # packed longest first
    assert list(rs) == list(np.concatenate([[0], np.cumsum(ln)[:-1]]))
    for v, r in zip(vids, res):
        want = _StubNative._row_scores(v.visual, v.audio)
        assert torch.equal(torch.as_tensor(r.scores), want)
        tag = int(round(float(want[0]) * 1e4))
        assert r.picks.shape == (len(v.cps),) and np.all(r.seg_mean == tag) and np.all(r.picks == tag % 2)
        assert r.summary.shape == (v.n_frames,) and np.all(r.summary == tag % 251)


def test_score_videos_returns_scores_in_the_callers_order():
    lengths = [5, 12, 5, 30, 1]
    vids = _mixed_videos(lengths, seed0=5)
    model = AVBiLSTMModel(16, 4, 512)
    stub = _StubNative()
    model.native = lambda for_training=False: stub
    out = model.score_videos([(v.visual, v.audio) for v in vids])
    assert list(stub.calls[0][1]) == sorted(lengths, reverse=True)
    for v, s in zip(vids, out):
        assert torch.equal(s, _StubNative._row_scores(v.visual, v.audio))


def test_evaluate_lays_rows_out_longest_first_and_keeps_dataset_order(monkeypatch):
    """scripts.evaluate.evaluate and data.dataset.DeviceDataset: rows longest video first, descriptors (and so the
    per-video metrics) in dataset order; the means equal the reference loop's (scripts/evaluate.py:12-42) on the
    same per-video scores."""
    from oracle import av_oracle
    from avsum_b200 import runtime
    from avsum_b200.data.dataset import DeviceDataset
    from avsum_b200.scripts.evaluate import evaluate
    lengths = [23, 64, 23, 9, 51]
    vids = _mixed_videos(lengths, seed0=900)
    rng = np.random.default_rng(3)
    dataset = [({"visual": v.visual, "audio": v.audio}, torch.as_tensor(rng.random(v.T).astype(np.float32)))
               for v in vids]

    def metrics_on_host(pred, target, row_start, lengths_):
        rows = []
        for s, n in zip(np.asarray(row_start), np.asarray(lengths_)):
            p, t = pred[s:s + n].numpy(), target[s:s + n].numpy()
            rows.append(list(av_oracle.eval_metrics(p, t)) + [float(np.mean(p))])
        return np.asarray(rows, np.float64), np.zeros((len(rows), 8), np.int64)

    monkeypatch.setattr(runtime, "eval_metrics_rows", metrics_on_host)
    want = [av_oracle.eval_metrics(_StubNative._row_scores(f["visual"], f["audio"]).numpy(), t.numpy())
            for f, t in dataset]
    model = _StubModel()
    out, per_video, _ = evaluate(model, dataset, return_per_video=True)
    rs, ln, _axis = model.nat.calls[0]
    assert list(ln) == lengths                                           # descriptors stay in dataset order
    assert list(rs[np.argsort(-ln, kind="stable")]) == list(np.concatenate([[0], np.cumsum(sorted(lengths, reverse=True))[:-1]]))
    assert np.array_equal(per_video[:, :3], np.asarray(want))
    assert out["f1"] == np.mean([w[0] for w in want]) and out["spearman"] == np.mean([w[1] for w in want])
    assert out["kendall"] == np.mean(np.asarray([w[2] for w in want], np.float32))

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)        # DeviceDataset refuses without CUDA
    dd = DeviceDataset(dataset, device="cpu")
    assert len(dd) == len(dataset) and list(dd.lengths) == lengths
    for i, (f, t) in enumerate(dataset):                                  # items come back in dataset order
        assert torch.equal(dd[i][0]["visual"], f["visual"]) and torch.equal(dd[i][0]["audio"], f["audio"])
        assert torch.equal(dd[i][1], t)
    assert dd.visual.shape[0] == sum(lengths) and torch.equal(dd.visual[:64], dataset[1][0]["visual"])   # longest first
    model2 = _StubModel()
    out2, per_video2, _ = evaluate(model2, dd, return_per_video=True)
    assert np.array_equal(per_video2, per_video) and out2 == out


# ---- the recurrence plan of the native library (avs_debug_plan: host logic of avs_forward, no GPU work) -------------

def _layout_longest_first(lengths):
    """row_start in the callers' layout: rows longest video first, descriptors in the caller's order."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    rs, at = np.zeros(len(lengths), np.int32), 0
    for i in order:
        rs[i] = at
        at += lengths[i]
    return rs


def test_recurrence_plan_of_config2_takes_the_per_group_schedule(native_lib):
    """config 2 = 7 groups of 8 slots, the extra video of the uneven split in the SHORTEST group,
    groups tiling the rows in order when the rows are laid out longest video first -- in the callers' order of the
    descriptors (evaluate / DeviceDataset) as well as in sorted order (packed_batches)."""
    from avsum_b200 import _cabi
    lengths = [v.T for v in synth.config2()]
    plan = _cabi.recurrence_plan(_layout_longest_first(lengths), lengths)
    assert plan["n_groups"] == 7 and plan["slots_per_group"] == 8
    assert plan["rows_ordered_by_group"] and plan["groups_end_apart"]
    sizes = np.bincount(plan["group_of"], minlength=7)
    assert list(sizes) == [7, 7, 7, 7, 7, 7, 8] and sizes.sum() == 50
    by_len = sorted(range(50), key=lambda i: -lengths[i])
    assert list(plan["group_of"][by_len]) == sorted(plan["group_of"])               # groups follow the length order
    rows = plan["group_rows"]
    assert rows[0, 0] == 0 and rows[-1, 1] == 21_477 and np.all(rows[1:, 0] == rows[:-1, 1])
    assert rows[0, 1] == sum(sorted(lengths, reverse=True)[:7]) == 4_522            # the 72 CTA-pair tiles of DESIGN 4
    # the same videos packed in the caller's (unsorted) order: same groups, but no group owns a block of rows
    unsorted = _cabi.recurrence_plan(np.concatenate([[0], np.cumsum(lengths)[:-1]]), lengths)
    assert np.array_equal(unsorted["group_of"], plan["group_of"]) and not unsorted["rows_ordered_by_group"]
    assert unsorted["group_rows"] is None


def test_recurrence_plan_edge_cases(native_lib):
    from avsum_b200 import _cabi
    # ties across a group boundary: the native sort is stable like the callers', so the layout still tiles
    lengths = [300] * 20 + [100] * 5
    plan = _cabi.recurrence_plan(_layout_longest_first(lengths), lengths)
    assert plan["n_groups"] == 4 and plan["rows_ordered_by_group"]
    assert not _cabi.recurrence_plan(_layout_longest_first([320] * 24), [320] * 24)["groups_end_apart"]
    # empty videos take no slot; padded layouts (rows no video owns) never tile
    lengths = [40, 0, 25, 0, 31]
    plan = _cabi.recurrence_plan(_layout_longest_first(lengths), lengths)
    assert plan["n_groups"] == 2                                  # <= 8 videos: two per cluster ...
    assert list(plan["group_of"]) == [0, -1, 1, -1, 1]            # ... the extra video in the SHORTEST group
    padded = _cabi.recurrence_plan([0, 50, 100], [40, 25, 31], total_rows=150)
    assert not padded["rows_ordered_by_group"]
    # small batches: 4 videos per cluster up to 16 videos, 8 for 17 - 64, wider clusters (one wave of 16) beyond
    assert _cabi.recurrence_plan(np.arange(16) * 10, [10] * 16)["n_groups"] == 4
    assert _cabi.recurrence_plan(np.arange(17) * 10, [10] * 17)["slots_per_group"] == 8
    assert _cabi.recurrence_plan(np.arange(64) * 10, [10] * 64)["slots_per_group"] == 8
    assert _cabi.recurrence_plan(np.arange(65) * 10, [10] * 65)["slots_per_group"] == 32
    empty = _cabi.recurrence_plan([], [])
    assert empty["n_groups"] == 0 and not empty["rows_ordered_by_group"]
    with pytest.raises(ValueError, match="outside"):
        _cabi.recurrence_plan([0, 5], [10, 10], total_rows=12)


def test_callers_layout_matches_the_native_plan(native_lib):
    """What summarize_videos / evaluate hand to the native call is a layout the per-group schedule accepts."""
    from avsum_b200 import _cabi
    from avsum_b200.evaluation.summary import summarize_videos
    rng = np.random.default_rng(11)
    lengths = [int(t) for t in rng.integers(20, 90, 30)]
    vids = _mixed_videos(lengths, seed0=3000)
    model = _StubModel()
    summarize_videos(model, vids)
    rs, ln, _ = model.nat.calls[0]
    plan = _cabi.recurrence_plan(rs, ln)
    assert plan["rows_ordered_by_group"] and plan["n_groups"] == 4 and plan["slots_per_group"] == 8
