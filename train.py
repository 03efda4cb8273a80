#!/usr/bin/env python
"""train.py — drop-in for the reference's training entry point (/root/reference/train.py:32-208).

Same flags, same per-epoch flow (train -> CosineAnnealingLR step -> evaluate -> best / latest checkpoint with the
reference's file names, train.py:127-149), same state-dict files.  The step itself (train.py:66-86) runs as
`sam2_unet_b200.TrainStep`: forward, the three structure_loss terms, backward and AdamW on the sm_100a kernels,
captured in one CUDA graph, with the loss read back only when it is printed (the reference syncs every step at
train.py:81).  Under torchrun the batch is sharded over ranks and gradients are all-reduced over NCCL.

Data: files are decoded to uint8 on the host workers; the reference's training transform (dataset.py:288-313) runs on
the device (`sam2_unet_b200.TrainAugment`, random decisions drawn in the reference's order), the evaluation set goes
through `preprocess_image` / `infer_tail` and the on-device metrics.  `--cpu_augment` uses the reference's own
`dataset.FullDataset` unchanged when it is importable; `--synthetic N` trains on N seeded synthetic images, no files.
"""
from __future__ import annotations

import argparse
import os
import time

import torch
import torch.distributed as dist

from sam2_unet_b200 import (SAM2UNet, TrainStep, cosine_lr, evaluate_dataset, evaluate_segmentation_performance,
                            print_eval_report,
                            infer_tail, preprocess_image)


def structure_loss(pred, mask):
    """train.py:21-29 — kept importable from this module like in the reference."""
    from sam2_unet_b200 import structure_loss as _sl
    return _sl(pred, mask)


class _RawFiles(torch.utils.data.Dataset):
    """The training FILES of the reference's FullDataset (dataset.py:288-333: sorted .jpg/.png images, .png masks), decoded
    to uint8 on the host workers; the transform itself (ToTensor ... Normalize, dataset.py:300-310) runs on the device in
    `sam2_unet_b200.TrainAugment`, with the random decisions drawn like the reference draws them."""

    def __init__(self, image_root, gt_root):
        self.images = sorted(os.path.join(image_root, f) for f in os.listdir(image_root) if f.endswith((".jpg", ".png")))
        self.gts = sorted(os.path.join(gt_root, f) for f in os.listdir(gt_root) if f.endswith(".png"))
        if len(self.images) != len(self.gts):
            raise ValueError(f"{len(self.images)} training images but {len(self.gts)} masks")

    def __len__(self):
        return len(self.images)

    def __getitem__(self, i):
        import numpy as np
        from PIL import Image
        img = np.array(Image.open(self.images[i]).convert("RGB"))
        gt = np.array(Image.open(self.gts[i]).convert("L"))
        return torch.from_numpy(img), torch.from_numpy(gt)


class _DeviceAugmentLoader:
    """Wraps a DataLoader of raw uint8 (image, mask) pairs: every batch is augmented on the GPU and handed on as the
    {"image": [B,3,S,S], "label": [B,1,S,S]} dictionary the reference's loader yields.  The augmentation kernels run on
    their own stream, so the batch for step i+1 (requested right after step i was launched) overlaps step i."""

    def __init__(self, loader, size, device):
        from sam2_unet_b200 import TrainAugment
        self.loader, self.aug, self.device = loader, TrainAugment(size, device), device
        self.stream = torch.cuda.Stream(device=device)

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for raw in self.loader:
            main = torch.cuda.current_stream(self.device)
            with torch.cuda.stream(self.stream):
                items = [self.aug(img, gt) for img, gt in raw]
                batch = {"image": torch.stack([d["image"] for d in items]),
                         "label": torch.stack([d["label"] for d in items])}
            main.wait_stream(self.stream)                    # the consumer's stream sees the finished batch
            for t in batch.values():
                t.record_stream(main)                        # allocated on the side stream, used on the main one
            yield batch


class _Synthetic(torch.utils.data.Dataset):
    def __init__(self, n, size):
        g = torch.Generator().manual_seed(0)
        yy, xx = torch.meshgrid(torch.arange(size, dtype=torch.float32), torch.arange(size, dtype=torch.float32), indexing="ij")
        self.items = []
        for _ in range(n):
            m = torch.zeros(size, size)
            for _ in range(3):
                cy, cx = (torch.rand(2, generator=g) * size).tolist()
                r = (0.08 + 0.2 * torch.rand(1, generator=g).item()) * size
                m = torch.maximum(m, ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).float())
            x = 0.5 * torch.randn(3, size, size, generator=g) + 2 * m - 0.5
            self.items.append({"image": x, "label": m[None]})

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


class _TestFiles:
    """The evaluation set of the reference (TestDataset, dataset.py:405-447): sorted image / ground-truth FILES, no
    augmentation.  Every item is decoded on the host and goes through the deterministic test-time transforms on the
    device (`preprocess_image` = dataset.py:336-407), exactly like test.py."""

    def __init__(self, image_root, gt_root):
        self.images = sorted(os.path.join(image_root, f) for f in os.listdir(image_root) if f.endswith((".jpg", ".png")))
        self.gts = sorted(os.path.join(gt_root, f) for f in os.listdir(gt_root) if f.endswith(".png"))
        if len(self.images) != len(self.gts):
            raise ValueError(f"{len(self.images)} test images but {len(self.gts)} ground-truth masks")

    def __len__(self):
        return len(self.images)

    def __getitem__(self, i):
        import numpy as np
        from PIL import Image
        img = np.array(Image.open(self.images[i]).convert("RGB"))
        gt = np.array(Image.open(self.gts[i]).convert("L"))
        return torch.from_numpy(img), torch.from_numpy(gt)


def _datasets(args):
    if args.synthetic > 0:
        ds = _Synthetic(args.synthetic, args.size)
        return ds, ds
    test = _TestFiles(args.test_image_path, args.test_gt_path)
    if args.cpu_augment:                             # the reference's own CPU pipeline, unchanged (needs dataset.py)
        from dataset import FullDataset
        return FullDataset(args.train_image_path, args.train_mask_path, args.size, mode="train"), test
    return _RawFiles(args.train_image_path, args.train_mask_path), test


@torch.no_grad()
def evaluate(model, dataset, device, size, rank=0, world=1):
    """Per-epoch evaluation of the reference (train.py:90-125), image by image: test-time transforms, forward, padding
    removed, bilinear resize to the ground-truth size, sigmoid, min-max, uint8, eval.py's per-image metrics - the pixel
    work of all of it on the device.  Under DDP every rank scores its stride of the file list; -> list of per-image
    result dictionaries of this rank (aggregate with evaluate_dataset after gathering)."""
    model.eval()
    results = []
    for i in range(rank, len(dataset), world):
        item = dataset[i]
        if isinstance(item, dict):                   # --synthetic: already at the network size, no padding
            res, _, _ = model(item["image"][None].to(device))
            gt = (item["label"][0].to(device) * 255).round().clamp(0, 255).to(torch.uint8)
            png = infer_tail(res, (0, 0, 0, 0), tuple(gt.shape))
        else:
            img_u8, gt = item
            x, padding = preprocess_image(img_u8.to(device), size)
            res, _, _ = model(x)
            gt = gt.to(device)
            png = infer_tail(res, padding, tuple(gt.shape))           # train.py:103-112 on the device
        results.append(evaluate_segmentation_performance(png, gt))
        if i % 10 == 0 and rank == 0:
            print(".", end="", flush=True)
    return results


def main(args):
    ddp = "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ddp:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if ddp else 0
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    train_ds, test_ds = _datasets(args)
    sampler = torch.utils.data.distributed.DistributedSampler(train_ds, shuffle=True) if ddp else None
    raw = isinstance(train_ds, _RawFiles)
    loader = torch.utils.data.DataLoader(train_ds, batch_size=args.batch_size, shuffle=sampler is None, sampler=sampler,
                                         num_workers=0 if args.synthetic else 8, drop_last=ddp, pin_memory=not raw,
                                         collate_fn=(lambda items: items) if raw else None)
    if raw:                                          # dataset.py:300-310 on the device (csrc/augment.cu)
        loader = _DeviceAugmentLoader(loader, args.size, device)
    if os.path.exists(args.hiera_path):
        hiera = args.hiera_path
    elif args.synthetic > 0 or args.random_trunk or len(args.checkpoint) > 0:
        hiera = ""                                   # smoke runs / full SAM2-UNet checkpoint: no pretrained trunk needed
        if rank == 0 and not len(args.checkpoint):
            print(f"WARNING: {args.hiera_path} not found - the frozen Hiera trunk keeps its RANDOM initialisation")
    else:                                            # the reference fails in build_sam2 (build_sam.py:79-89): so do we
        raise FileNotFoundError(f"--hiera_path {args.hiera_path} does not exist: the trunk is frozen, training on a random "
                                "encoder is never what you want (pass --random_trunk to do it anyway)")
    model = SAM2UNet(checkpoint_path=hiera, model_cfg=args.model_cfg, dtype=args.dtype).to(device)
    if len(args.checkpoint) > 0:
        model.load_state_dict(torch.load(args.checkpoint, map_location=device), strict=True)
    if ddp:                                          # one set of initial weights / BN buffers: rank 0's
        model._engine(device)
        dist.broadcast(model.flat.master, 0)
        for b in model.buffers():
            dist.broadcast(b, 0)
        model.flat.bump()
    step = TrainStep(model, lr=args.lr, weight_decay=args.weight_decay, use_graph=not args.no_graph)
    os.makedirs(args.save_path, exist_ok=True)
    log_path = os.path.join(args.save_path, "log.txt")
    base_mean_iou = args.base_mean_iou
    for epoch in range(args.epoch):
        if rank == 0:
            print("Training:")
        model.train()
        if sampler is not None:
            sampler.set_epoch(epoch)
        step.optim.param_groups[0]["lr"] = cosine_lr(epoch, args.epoch, args.lr)      # train.py:54,87
        t0, seen, loss = time.time(), 0, None
        it = iter(loader)
        batch, i = next(it, None), 0
        while batch is not None:
            loss = step(batch["image"], batch["label"])
            nxt = next(it, None)                             # the next (pinned) batch starts its H2D copy while this
            if nxt is not None:                              # step's kernels run
                step.prefetch(nxt["image"], nxt["label"])
            seen += batch["image"].shape[0]
            if i % 10 == 0 and rank == 0:
                print("epoch-{}-{}: loss:{}".format(epoch + 1, i + 1, loss.sum().item()))
            batch, i = nxt, i + 1
        torch.cuda.synchronize(device)
        epoch_loss = loss.sum().item() if loss is not None else float("nan")
        if rank == 0:
            print(f"epoch {epoch + 1}: {seen / max(time.time() - t0, 1e-9):.1f} img/s/rank")
            print("Evaluating", end="")
        # every rank scores its share of the test files (no rank idles in an NCCL barrier while rank 0 evaluates)
        mine = evaluate(model, test_ds, device, args.size, rank, dist.get_world_size() if ddp else 1)
        if ddp:
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, mine)
            mine = [r for part in parts for r in part]
        if rank == 0:
            final = evaluate_dataset(mine)
            mean_iou = final.get("mIoU", 0.0)
            epoch_name = f"epoch-{epoch + 1}_loss-{epoch_loss:.3f}"
            print()
            print_eval_report(final, title=epoch_name, log_path=log_path)           # train.py:126-128, same layout
            if mean_iou > base_mean_iou:                                              # train.py:133-143
                base_mean_iou = mean_iou
                path = os.path.join(args.save_path, f"SAM2-UNet_{epoch_name}_iou-{mean_iou:.3f}.pth")
                torch.save(model.state_dict(), path)
                print("Saving Snapshot best:", path)
            elif (epoch + 1) % args.save_interval == 0 or (epoch + 1) == args.epoch:  # train.py:144-149
                path = os.path.join(args.save_path, "SAM2-UNet_epoch-latest.pth")
                torch.save(model.state_dict(), path)
                print("Saving Snapshot:", path)
        if ddp:
            dist.barrier()
    if ddp:
        dist.destroy_process_group()


if __name__ == "__main__":
    parser = argparse.ArgumentParser("SAM2-UNet")
    parser.add_argument("--save_path", type=str, required=True, help="path to store the checkpoint")
    parser.add_argument("--hiera_path", type=str, default="../sam2_hiera_small.pt", help="path to the sam2 pretrained hiera")
    parser.add_argument("--checkpoint", type=str, default="", help="path to the checkpoint of sam2-unet")
    parser.add_argument("--train_image_path", type=str, default="data_train/images/")
    parser.add_argument("--train_mask_path", type=str, default="data_train/masks/")
    parser.add_argument("--test_image_path", type=str, default="data_test/images/")
    parser.add_argument("--test_gt_path", type=str, default="data_test/masks/")
    parser.add_argument("--epoch", type=int, default=500, help="training epochs")
    parser.add_argument("--lr", type=float, default=0.001, help="learning rate")
    parser.add_argument("--batch_size", default=16, type=int)
    parser.add_argument("--size", default=960, type=int)
    parser.add_argument("--weight_decay", default=5e-4, type=float)
    parser.add_argument("--save_interval", default=20, type=int)
    parser.add_argument("--base_mean_iou", default=0.83, type=float)
    # extras of this implementation
    parser.add_argument("--model_cfg", default="sam2_hiera_s.yaml", help="trunk yaml: sam2_hiera_{t,s,b+,l}.yaml")
    parser.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    parser.add_argument("--no_graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    parser.add_argument("--cpu_augment", action="store_true",
                        help="augment on the host with the reference's dataset.py instead of the device pipeline")
    parser.add_argument("--synthetic", type=int, default=0, help="train on N seeded synthetic images (no files needed)")
    parser.add_argument("--random_trunk", action="store_true", help="allow training without the pretrained Hiera weights")
    main(parser.parse_args())
