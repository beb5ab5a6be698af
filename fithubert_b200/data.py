"""Data side of the distillation loop: the reference's length-bucketed LibriSpeech dataset
(utils/dataset.py:11-77) and what Lightning's DataLoader + DistributedSampler did around it (train.py:423-441,494),
without pandas-in-the-hot-loop or Lightning: buckets are decoded, padded and PINNED by a background thread while the
GPU works on the previous one, and every data-parallel rank walks its own slice of the bucket list.

Same input contract as the reference: a bucket is `{'x': fp32 [B, Lmax] zero padded, 'padding_mask': bool [B, Lmax]}`
(True = padding), keys matching `W2V2Distil.forward(x, padding_mask)` / `training_step(batch)`.
"""
from __future__ import annotations

import os
import queue
import random
import threading
import wave as _wave
from typing import Dict, Iterator, List, Optional, Sequence

import torch
from torch.nn.utils.rnn import pad_sequence


def load_audio(path: str) -> torch.Tensor:
    """1-D fp32 waveform in [-1, 1].  .flac / anything else through torchaudio or soundfile when the image has a
    decoder; 16-bit PCM .wav through the standard library (what the tests use - no codec needed)."""
    if path.lower().endswith(".wav"):
        try:
            with _wave.open(path, "rb") as f:
                if f.getsampwidth() == 2:
                    raw = f.readframes(f.getnframes())
                    x = torch.frombuffer(bytearray(raw), dtype=torch.int16).float() / 32768.0
                    return x.view(-1, f.getnchannels())[:, 0].contiguous()
        except _wave.Error:
            pass
    try:
        import torchaudio
        wav, _ = torchaudio.load(path)  # reference utils/dataset.py:59-61
        return wav.squeeze()
    except Exception as first:  # noqa: BLE001 - torchaudio without a backend raises assorted types
        try:
            import soundfile
            data, _ = soundfile.read(path, dtype="float32")
            x = torch.from_numpy(data)
            return x[:, 0].contiguous() if x.dim() > 1 else x
        except ImportError:
            raise RuntimeError(f"no audio decoder available for {path}: {first}") from first


class LibriDataset(torch.utils.data.Dataset):
    """Librispeech waveform dataset, bucketed by length (reference utils/dataset.py:11-77).

    `file_path/<set>.csv` must carry the columns `file_path` and `length` (the s3prl `len_for_bucket` tables).
    Utterances are sorted by length (descending) and cut into consecutive buckets of `batch_size`; a trailing
    bucket with a single utterance is dropped (:50-53).  Index = one whole bucket (the DataLoader runs with
    batch_size = 1, train.py:424-428)."""

    def __init__(self, batch_size, file_path="/workspace/s3prl/s3prl/data/len_for_bucket/",
                 sets=("train-clean-100", "train-clean-360", "train-other-500"), libri_root="/workspace/LibriSpeech/"):
        super().__init__()
        import pandas as pd
        self.libri_root = libri_root
        self.root = file_path
        tables = [pd.read_csv(os.path.join(file_path, s + ".csv")) for s in sets]
        self.table = pd.concat(tables, ignore_index=True).sort_values(by=["length"], ascending=False)
        X = self.table["file_path"].tolist()
        X_lens = self.table["length"].tolist()
        self.num_samples = len(X)
        self.X: List[List[str]] = []
        self.bucket_lengths: List[List[int]] = []
        batch_x, batch_len = [], []
        for x, x_len in zip(X, X_lens):
            batch_x.append(x)
            batch_len.append(int(x_len))
            if len(batch_x) == batch_size:
                self.X.append(batch_x)
                self.bucket_lengths.append(batch_len)
                batch_x, batch_len = [], []
        if len(batch_x) > 1:
            self.X.append(batch_x)
            self.bucket_lengths.append(batch_len)

    def collate_fn(self, items):
        return items[0]

    def _load_feat(self, feat_path):
        return load_audio(os.path.join(self.libri_root, feat_path))

    def __getitem__(self, index) -> Dict[str, torch.Tensor]:
        wave_orig = [self._load_feat(x_file) for x_file in self.X[index]]
        wav_lengths = torch.LongTensor([len(wav) for wav in wave_orig])
        wav_padding_mask = ~torch.lt(torch.arange(int(max(wav_lengths))).unsqueeze(0), wav_lengths.unsqueeze(1))
        padded_wav = pad_sequence(wave_orig, batch_first=True)
        return {"x": padded_wav, "padding_mask": wav_padding_mask}

    def __len__(self):
        return len(self.X)


class SyntheticBuckets(torch.utils.data.Dataset):
    """Bucketed synthetic waveforms with the LibriSpeech-like length distribution of BASELINE.md section 4 (for
    benchmarks and tests: there is no dataset in the image).  Same item contract as LibriDataset."""

    def __init__(self, n_buckets: int, batch_size: int, max_len: int, min_len: Optional[int] = None, seed: int = 0):
        g = random.Random(seed)
        lo = min_len if min_len is not None else max(400, max_len - 8000)
        lens = sorted((g.randint(lo, max_len) for _ in range(n_buckets * batch_size)), reverse=True)
        self.bucket_lengths = [lens[i:i + batch_size] for i in range(0, len(lens), batch_size)]
        self.seed = seed

    def __len__(self):
        return len(self.bucket_lengths)

    def collate_fn(self, items):
        return items[0]

    def __getitem__(self, index):
        lens = self.bucket_lengths[index]
        gen = torch.Generator().manual_seed(self.seed * 1000003 + index)
        x = 0.1 * torch.randn(len(lens), lens[0], generator=gen)
        pm = ~(torch.arange(lens[0]).unsqueeze(0) < torch.tensor(lens).unsqueeze(1))
        return {"x": x.masked_fill_(pm, 0.0), "padding_mask": pm}


def shard_indices(n: int, rank: int, world: int, shuffle: bool, seed: int, epoch: int) -> List[int]:
    """torch DistributedSampler semantics (what Lightning wraps the reference's DataLoader in under strategy='ddp'):
    a permutation seeded by seed + epoch, padded by wrap-around to a multiple of `world`, then rank::world."""
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g).tolist()
    else:
        idx = list(range(n))
    if world > 1:
        total = -(-n // world) * world
        pad = total - n
        if pad > 0 and n > 0:  # DistributedSampler repeats the list when the padding is longer than the list itself
            idx += (idx * -(-pad // n))[:pad]
        idx = idx[rank:total:world]
    return idx


class BucketLoader:
    """Iterates one rank's buckets of an epoch; a background thread decodes / pads the next `prefetch` buckets into
    PINNED host memory so that the step's host->device copy is a plain async DMA (replaces DataLoader(batch_size=1,
    shuffle=True, collate_fn=..., num_workers=4*gpus), train.py:423-428)."""

    def __init__(self, dataset, shuffle: bool = True, rank: int = 0, world: int = 1, seed: int = 0, prefetch: int = 2,
                 pin: Optional[bool] = None):
        self.dataset, self.shuffle, self.rank, self.world, self.seed = dataset, shuffle, rank, world, seed
        self.prefetch = max(1, prefetch)
        self.pin = torch.cuda.is_available() if pin is None else pin
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __len__(self) -> int:
        return len(shard_indices(len(self.dataset), self.rank, self.world, False, 0, 0))

    @staticmethod
    def _put(out: "queue.Queue", item, stop: "threading.Event") -> bool:
        """Blocking put that gives up once the consumer has gone away (early break / exception in the training loop)."""
        while not stop.is_set():
            try:
                out.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _produce(self, order: Sequence[int], out: "queue.Queue", stop: "threading.Event") -> None:
        try:
            for i in order:
                if stop.is_set():
                    return
                item = self.dataset[i]
                if self.pin:
                    item = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in item.items()}
                if not self._put(out, item, stop):
                    return
            self._put(out, None, stop)
        except BaseException as e:  # noqa: BLE001 - handed to the consumer
            self._put(out, e, stop)

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        order = shard_indices(len(self.dataset), self.rank, self.world, self.shuffle, self.seed, self.epoch)
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()
        t = threading.Thread(target=self._produce, args=(order, q, stop), daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:  # also reached when the consumer stops early (break, exception, generator close)
            stop.set()
            t.join()
