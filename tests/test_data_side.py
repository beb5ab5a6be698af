"""CPU-only: the data side - length-bucketed dataset, per-rank sharding, prefetching loader and the
Trainer pieces (schedule length, top-k checkpoints, early stopping) against the reference's logic restated inline
(utils/dataset.py:27-74, train.py:411-413,475-490) and torch's own DistributedSampler."""
import os
import struct
import wave

import pytest
import torch

import fithubert_b200 as F
from fithubert_b200 import data as D, trainer as TR


def _write_wav(path, n, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, generator=g) * 2 - 1).mul(20000).short()
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(16000)
        f.writeframes(struct.pack(f"<{n}h", *x.tolist()))
    return x.float() / 32768.0


@pytest.fixture()
def libri(tmp_path):
    root = tmp_path / "LibriSpeech"
    os.makedirs(root / "a")
    lens = {"a/u0.wav": 900, "a/u1.wav": 1500, "a/u2.wav": 1200, "a/u3.wav": 700, "a/u4.wav": 1501,
            "a/u5.wav": 640, "a/u6.wav": 1000}
    waves = {k: _write_wav(str(root / k), n, i) for i, (k, n) in enumerate(lens.items())}
    bucket_dir = tmp_path / "len_for_bucket"
    os.makedirs(bucket_dir)
    items = list(lens.items())
    for name, part in (("train-a", items[:4]), ("train-b", items[4:])):
        with open(bucket_dir / f"{name}.csv", "w") as f:
            f.write(",file_path,length,label\n")
            for i, (k, n) in enumerate(part):
                f.write(f"{i},{k},{n},x\n")
    return str(bucket_dir), str(root), lens, waves


def test_libri_dataset_buckets_like_the_reference(libri):
    bucket_dir, root, lens, waves = libri
    ds = F.LibriDataset(batch_size=3, file_path=bucket_dir, sets=["train-a", "train-b"], libri_root=root)
    # reference logic restated (utils/dataset.py:27-53): sort by length descending, consecutive buckets, a trailing
    # bucket is kept only if it has more than one utterance
    order = sorted(lens, key=lambda k: -lens[k])
    ref = [order[i:i + 3] for i in range(0, len(order), 3)]
    ref = [b for b in ref if len(b) == 3 or len(b) > 1]
    assert ds.X == ref and len(ds) == 2 and ds.num_samples == 7  # the 7th utterance alone is dropped
    item = ds.collate_fn([ds[1]])
    ls = [lens[k] for k in ds.X[1]]
    assert item["x"].shape == (3, max(ls)) and item["x"].dtype == torch.float32 and item["padding_mask"].dtype == torch.bool
    for r, k in enumerate(ds.X[1]):
        assert torch.equal(item["x"][r, :ls[r]], waves[k]) and float(item["x"][r, ls[r]:].abs().sum()) == 0.0
        assert item["padding_mask"][r].tolist() == [False] * ls[r] + [True] * (max(ls) - ls[r])
    ds4 = F.LibriDataset(batch_size=4, file_path=bucket_dir, sets=["train-a", "train-b"], libri_root=root)
    assert [len(b) for b in ds4.X] == [4, 3]


@pytest.mark.parametrize("n,world", [(10, 1), (10, 2), (11, 4), (7, 8)])
def test_shard_indices_equal_distributed_sampler(n, world):
    from torch.utils.data import DistributedSampler
    ds = list(range(n))
    for epoch in (0, 3):
        for shuffle in (False, True):
            seen = []
            for rank in range(world):
                s = DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=shuffle, seed=5)
                s.set_epoch(epoch)
                mine = F.shard_indices(n, rank, world, shuffle, 5, epoch)
                assert mine == list(iter(s))
                seen += mine
            assert set(seen) == set(range(n))


def test_bucket_loader_prefetches_in_order_and_propagates_errors():
    ds = F.SyntheticBuckets(n_buckets=5, batch_size=2, max_len=3000, seed=1)
    got = []
    for rank in range(2):
        ld = F.BucketLoader(ds, shuffle=True, rank=rank, world=2, seed=9, pin=False)
        ld.set_epoch(2)
        batches = list(ld)
        assert len(batches) == len(ld) == 3
        order = F.shard_indices(5, rank, 2, True, 9, 2)
        for b, i in zip(batches, order):
            ref = ds[i]
            assert torch.equal(b["x"], ref["x"]) and torch.equal(b["padding_mask"], ref["padding_mask"])
            assert b["x"].shape[1] == ds.bucket_lengths[i][0] and not bool(b["padding_mask"][0].any())
        got += order
    assert set(got) == set(range(5))

    class Broken(D.SyntheticBuckets):
        def __getitem__(self, i):
            if i == 1:
                raise OSError("corrupt file")
            return super().__getitem__(i)

    with pytest.raises(OSError):
        list(F.BucketLoader(Broken(3, 2, 1000), shuffle=False, pin=False))


def test_trainer_pieces(tmp_path):
    # train.py:411-412 with fithubert.yaml on 4 GPUs: 100 epochs, accumulate 4
    assert F.total_training_steps(8793, 4, 100, 4) == (100 * (8793 // 4)) // 4
    ck = TR.TopKCheckpoints(str(tmp_path), k=3)
    for epoch, v in enumerate([5.0, 4.0, 6.0, 3.0, 7.0, 3.5]):
        ck.save({"epoch": epoch, "state_dict": {}}, epoch, v)
    kept = sorted(f for f in os.listdir(tmp_path) if f.startswith("checkpoint-"))
    assert kept == ["checkpoint-epoch=01.ckpt", "checkpoint-epoch=03.ckpt", "checkpoint-epoch=05.ckpt"]
    assert torch.load(os.path.join(tmp_path, "last.ckpt"))["epoch"] == 5
    es = TR.EarlyStopping(patience=3)
    assert [es.step(v) for v in (1.0, 0.9, 0.95, 0.91, 0.93)] == [False, False, False, False, True]
