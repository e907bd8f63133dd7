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
