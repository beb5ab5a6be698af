"""`python -m fithubert_b200.train --config data/conf/fithubert.yaml` - the reference's train.py entry point
(train.py:452-509) on the B200 path: one process per GPU (launch with torchrun for gpus > 1; Lightning's
`strategy='ddp'` spawned the ranks itself), the yaml drives everything exactly as in the reference."""
from __future__ import annotations

import argparse
import os

import torch
import torch.distributed as dist
import yaml


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-c", "-cfg", "--config", help="yaml config path for training")
    parser.add_argument("-t", "--test", action="store_true", help="Enable testing mode")
    args = parser.parse_args(argv)
    yaml_path = args.config or "./data/conf/ex.yaml"
    with open(yaml_path) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)
    t = cfg["train"]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        # the gradient all-reduce runs under the backward: NCCL is capped at the SMs the persistent kernels leave it
        # (W2V2Distil.fused_forward_backward, FHB_COMM_SMS) - 125 MB over NVLink need no more
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("FHB_COMM_SMS", "8"))
        dist.init_process_group("nccl", device_id=dev)
    from . import BucketLoader, LibriDataset, W2V2Distil, fit
    model = W2V2Distil(cfg, device=dev)
    data_cfg = cfg["data"]
    mk = lambda sets: LibriDataset(batch_size=t["batch_size"], file_path=data_cfg["bucketing_path"], sets=sets,  # noqa: E731
                                   libri_root=data_cfg["libri_root"])
    output_dir = "./results/pretrain/" + t["output_dir"]
    if args.test:
        loader = BucketLoader(mk(data_cfg["test_set"]), shuffle=False, rank=rank, world=world)
        model.student_model.eval()
        tot, cnt = 0.0, 0
        for i, batch in enumerate(loader):
            tot += float(model.validation_step(batch, i)["v_loss"]) * batch["x"].shape[0]
            cnt += batch["x"].shape[0]
        print({"test_loss": tot / max(1, cnt)})
        return
    train = BucketLoader(mk(data_cfg["train_set"]), shuffle=True, rank=rank, world=world)
    val = BucketLoader(mk(["dev-clean"]), shuffle=False, rank=rank, world=world)
    ckpt = os.path.join(output_dir, t["checkpoint"]) if t.get("checkpoint") else None
    res = fit(model, train, val, num_epochs=t["num_epochs"], output_dir=output_dir, ckpt_path=ckpt, log_every=50)
    if rank == 0:
        print(res)


if __name__ == "__main__":
    main()
