"""N > 1 host logic on CPU: world_size-2 gloo run of the bucketed gradient all-reduce (optim.GradAllReduce)
and the DDP semantics the bench relies on (per-rank mean loss, gradient average across ranks ==
single-process gradient of the mean over the concatenated batch, for equal-shape shards)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fhb_oracle as O
    from fithubert_b200.optim import GradAllReduce
    torch.manual_seed(0)
    # a tiny student, identical on every rank; each rank gets its own shard of a 2*world batch
    scfg = O.student_config(conv_feature_layers="[(16,10,5)] + [(32,1,1)] + [(32,2,2)] * 2", encoder_layers=1,
                            encoder_embed_dim=32, encoder_ffn_embed_dim=32, encoder_attention_heads=2, conv_pos=8,
                            conv_pos_groups=2, pred_head_final_dim=16)
    sd = O.init_student_state(scfg, 0, perturb=True)
    x_all, pm_all = O.synth_batch(2 * world, 2000, [2000, 1500] * world, seed=3)
    tgt_all = torch.randn(2 * world, 1, 99, 16, generator=torch.Generator().manual_seed(5))

    def grads(x, pm, tgt):
        p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        s = O.student_forward(p, scfg, x, pm)
        pred = torch.stack(s["projections"], 1)
        loss = ((pred - tgt[:, :, :pred.shape[2]]) ** 2).mean()
        loss.backward()
        names = [k for k in p if p[k].grad is not None]
        return names, torch.cat([p[k].grad.flatten() for k in names]), float(loss)

    names, flat, loss = grads(x_all[2 * rank:2 * rank + 2], pm_all[2 * rank:2 * rank + 2], tgt_all[2 * rank:2 * rank + 2])
    red = GradAllReduce(n_buckets=5)
    assert red.enabled and red.world == world and red.stream is None
    bounds = red.bucket_bounds(flat.numel())
    assert bounds[0][0] == 0 and bounds[-1][1] == flat.numel()
    assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
    split = flat.clone()
    red.reduce_all(flat)
    red.wait()
    # two-phase exchange (reduce_tail as soon as the layers' / heads' gradients are final, the rest after the backward):
    # every element is reduced exactly once, whatever the split point
    cut = (flat.numel() // 3) + 1
    red.reduce_tail(split, cut)
    red.reduce_all(split)
    red.wait()
    assert torch.equal(split, flat)
    red.reduce_all(split)  # the next step starts from a clean state: a full exchange again
    assert torch.allclose(split, flat * world)
    flat = flat / red.world  # the AdamW kernel's grad_scale = 1/world
    _, ref, _ = grads(x_all, pm_all, tgt_all)
    err = float((flat - ref).abs().max() / ref.abs().max())
    q.put((rank, err, len(bounds)))
    dist.destroy_process_group()


def test_two_rank_bucketed_allreduce_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, nb in res:
        assert err < 1e-5, (rank, err)
        assert nb == 5


def _fit_worker(rank, world, port, q):
    sys.path[:0] = [ROOT]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import types
    from fithubert_b200.trainer import fit

    class Shard:
        """Stands in for W2V2Distil on one rank (fit only touches this surface).  Rank 0's validation shard keeps improving,
        rank 1's gets worse: a rank-local EarlyStopping would fire on rank 1 alone and leave rank 0 blocked in the next
        collective; the global mean (1 + 0.02 * epoch) gets worse from the start, so BOTH must stop after `patience` epochs."""

        def __init__(self):
            self.train_cfg = {"num_epochs": 40}
            self.accumulate, self._micro, self.optimizer, self.epoch = 1, 0, None, 0
            self.student_model = types.SimpleNamespace(train=lambda *a: None, eval=lambda: None)

        def configure_optimizers(self, total_steps=0):
            self.total_steps = total_steps

        def training_step(self, batch, i):
            return torch.tensor(2.0 + rank)

        def training_epoch_end(self):
            self.epoch += 1

        def validation_step(self, batch, i):
            return {"v_loss": torch.tensor(1.0 - 0.01 * self.epoch if rank == 0 else 1.0 + 0.05 * self.epoch)}

    batches = [{"x": torch.zeros(2, 8)} for _ in range(3)]
    res = fit(Shard(), batches, batches, patience=3)
    q.put((rank, res["epochs_run"], res["stopped_early"], [round(h[2], 6) for h in res["history"]],
           [round(h[1], 6) for h in res["history"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_fit_stops_on_the_same_epoch_on_every_rank():
    """trainer.fit under world_size 2 (gloo): the monitored v_loss and the early-stop decision are reduced over the ranks
    (Lightning does both), so ranks whose shards disagree still leave the loop together."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fit_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, e0, s0, v0, t0), (_, e1, s1, v1, t1) = res
    assert e0 == e1 == 4 and s0 and s1          # best at epoch 0, then 3 epochs without improvement
    assert v0 == v1 and t0 == t1                # every rank saw the same (global) numbers
    assert abs(v0[0] - 1.02) < 1e-6 and abs(v0[1] - 1.04) < 1e-6 and abs(t0[0] - 2.5) < 1e-6  # global means
