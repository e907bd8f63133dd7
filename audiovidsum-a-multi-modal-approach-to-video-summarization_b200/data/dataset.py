"""On-disk feature contract of the reference (``data/dataset.py:8-32``; written by ``scripts/preprocess.py:74-81``):

    <feature_dir>/<video_id>/{visual.npy, audio.npy, scores.npy}      (float32 arrays, one row per sampled frame)

``BaseDataset`` keeps the reference's constructor and item type -- ``(features: {"visual", "audio"}, scores)`` as
torch tensors -- so ``scripts/evaluate.evaluate(model, dataset)`` works unchanged.  ``packed_batches`` is the
feeding side the reference lacks: it turns a dataset into length-bucketed, PACKED batches in pinned host
memory (one row per frame plus ``row_start`` / ``lengths`` descriptors), the layout ``avs_forward`` consumes, so
that the per-video ``.cuda()`` / ``.cpu()`` synchronisation of ``scripts/evaluate.py:13-15`` disappears.
"""
from __future__ import annotations

import os
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch


class BaseDataset(torch.utils.data.Dataset):
    def __init__(self, feature_dir, annotation_path=None):
        self.feature_dir = feature_dir
        self.video_ids = sorted(os.listdir(feature_dir))   # sorted: deterministic order (os.listdir is not)
        if annotation_path is not None:
            raise NotImplementedError(
                "annotation files are not part of the feature contract: the reference's BaseDataset._load_annotations "
                "is itself undefined (data/dataset.py:12-14); scores come from <video>/scores.npy")
        self.annotations = None

    def __len__(self):
        return len(self.video_ids)

    def __getitem__(self, idx):
        vid = self.video_ids[idx]
        features = {
            "visual": torch.from_numpy(np.load(os.path.join(self.feature_dir, vid, "visual.npy"))),
            "audio": torch.from_numpy(np.load(os.path.join(self.feature_dir, vid, "audio.npy"))),
        }
        scores = torch.from_numpy(np.load(os.path.join(self.feature_dir, vid, "scores.npy")))
        return features, scores


class PackedBatch:
    """Packed variable-length batch in pinned host memory."""

    def __init__(self, indices, visual, audio, scores, row_start, lengths):
        self.indices = indices          # dataset indices of the videos, in batch order
        self.visual = visual            # fp32 (or fp16, feature_dtype="fp16") [sum T, Dv] pinned
        self.audio = audio              # fp32 (or fp16) [sum T, Da] pinned
        self.scores = scores            # [sum T] pinned (dtype of scores.npy)
        self.row_start = row_start      # int32 [n]
        self.lengths = lengths          # int32 [n]


def packed_batches(dataset, max_frames: int = 32768, max_videos: int = 64, bucket: bool = True,
                   pin: Optional[bool] = None, feature_dtype: str = "fp32") -> Iterator[PackedBatch]:
    """Yield PackedBatch objects covering the dataset once.

    ``feature_dtype="fp16"`` (opt-in) packs the features as IEEE half: the batch is half as large on its way across
    PCIe -- the end-to-end step of a streamed evaluation is transfer-bound -- and the library reads it directly
    (``avs_model_set_feature_format``; the fc layers of the default precision mode consume 11-bit significands either
    way).  Values must fit the fp16 range (|x| < 65504; CNN / VGGish features do); the default stays float32, the
    reference's on-disk type (``data/dataset.py:25-28``).

    Videos are sorted by length (longest first) when ``bucket`` is set, so that the videos of one batch -- which
    share LSTM clusters whose run time is set by their longest member -- have similar lengths; a batch closes
    when adding a video would exceed ``max_frames`` rows or ``max_videos`` videos.
    """
    if feature_dtype not in ("fp32", "fp16"):
        raise ValueError("feature_dtype must be 'fp32' or 'fp16'")
    fdt = torch.float16 if feature_dtype == "fp16" else torch.float32
    pin = torch.cuda.is_available() if pin is None else pin
    items = [dataset[i] for i in range(len(dataset))]
    order = list(range(len(items)))
    if bucket:
        order.sort(key=lambda i: -int(items[i][0]["visual"].shape[0]))
    cur: List[int] = []
    rows = 0

    def flush(idxs):
        lens = np.asarray([int(items[i][0]["visual"].shape[0]) for i in idxs], dtype=np.int32)
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
        visual = torch.cat([items[i][0]["visual"].to(fdt) for i in idxs])
        audio = torch.cat([items[i][0]["audio"].to(fdt) for i in idxs])
        scores = torch.cat([items[i][1].reshape(-1) for i in idxs])
        if pin:
            visual, audio, scores = visual.pin_memory(), audio.pin_memory(), scores.pin_memory()
        return PackedBatch(list(idxs), visual, audio, scores, starts, lens)

    for i in order:
        t = int(items[i][0]["visual"].shape[0])
        if cur and (rows + t > max_frames or len(cur) >= max_videos):
            yield flush(cur)
            cur, rows = [], 0
        cur.append(i)
        rows += t
    if cur:
        yield flush(cur)


class DeviceDataset:
    """A whole dataset packed ONCE into device memory (``scripts/evaluate.py:12-18`` re-uploads every video on every
    call: ``.cuda()`` per video, ``:13-14``).  Evaluation loops that score the same dataset repeatedly -- every epoch
    of ``scripts/train_av_model.py``, hyper-parameter sweeps -- pass this object to ``scripts.evaluate.evaluate``
    instead of the ``BaseDataset``: the features cross PCIe once (4.6 KB per frame; 2.3 KB with
    ``feature_dtype="fp16"``) and every later evaluation is the device-resident step.

    Iterating yields the reference's item type ``({"visual", "audio"}, scores)`` (device tensor views), so code
    written against ``BaseDataset`` keeps working.
    """

    def __init__(self, dataset, device=None, feature_dtype: str = "fp32"):
        if feature_dtype not in ("fp32", "fp16"):
            raise ValueError("feature_dtype must be 'fp32' or 'fp16'")
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceDataset needs a CUDA device (avsum_b200 has no CPU fallback)")
        fdt = torch.float16 if feature_dtype == "fp16" else torch.float32
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        items = [dataset[i] for i in range(len(dataset))]
        self.lengths = np.asarray([int(f["visual"].shape[0]) for f, _ in items], dtype=np.int32)
        # rows are laid out longest video first (each recurrence group then owns one block of rows: the native forward
        # runs a group's tail behind its own recurrence); row_start / lengths stay indexed by dataset position
        layout = sorted(range(len(items)), key=lambda i: -int(self.lengths[i]))
        self.row_start = np.zeros(len(items), dtype=np.int32)
        at = 0
        for i in layout:
            self.row_start[i] = at
            at += int(self.lengths[i])
        if items:
            self.visual = torch.cat([torch.as_tensor(items[i][0]["visual"]).to(fdt) for i in layout]).to(dev)
            self.audio = torch.cat([torch.as_tensor(items[i][0]["audio"]).to(fdt) for i in layout]).to(dev)
            tgt = [torch.as_tensor(t).reshape(-1) for _, t in items]
            tdt = torch.float64 if any(t.dtype == torch.float64 for t in tgt) else torch.float32
            self.scores = torch.cat([tgt[i].to(tdt) for i in layout]).to(dev)
        else:
            self.visual = self.audio = self.scores = torch.zeros(0, device=dev)
        for n, (_, t) in zip(self.lengths, items):
            if int(torch.as_tensor(t).numel()) != int(n):
                raise ValueError("every video needs one target score per frame")

    def __len__(self):
        return int(self.lengths.size)

    def __getitem__(self, idx):
        s, n = int(self.row_start[idx]), int(self.lengths[idx])
        return {"visual": self.visual[s:s + n], "audio": self.audio[s:s + n]}, self.scores[s:s + n]

    def __iter__(self):
        return (self[i] for i in range(len(self)))
