"""Informative ("also report"): the UNMODIFIED reference modules in PyTorch eager on the same
B200 - fp32 and torch.autocast(bf16) - on the benchmark configuration (Hiera-L 352x352, batch 12, train step and
inference forward).  This is the real "library" bar: the reference ships no GPU kernels of its own, so on a GPU it runs
on cuDNN / cuBLAS / SDPA.  Not a bench arm; writes one JSON line per case.

    python scripts/ref_gpu_eager.py [--batch 12] [--size 352] [--steps 10] > gpurun_out/ref_gpu_eager.jsonl
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=12)
    ap.add_argument("--size", type=int, default=352)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--variant", default="l")
    args = ap.parse_args()
    for root in (os.environ.get("SAM2UNET_REFERENCE", ""), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if root and os.path.isfile(os.path.join(root, "SAM2UNet.py")):
            os.environ["SAM2UNET_REFERENCE"] = root
            break
    else:
        print(json.dumps({"unavailable": "reference files not staged (scripts/stage_reference.py)"}))
        return
    from oracle import ref_shim
    from sam2_unet_b200.params import fill_deterministic_
    from sam2_unet_b200.synthetic import synthetic_batch
    dev = torch.device("cuda", 0)
    x, mask = synthetic_batch(args.batch, args.size, seed=0)
    x, mask = x.to(dev), mask.to(dev)
    model = ref_shim.build_reference(args.variant)
    fill_deterministic_(model, 0)
    model.to(dev)
    loss_fn = ref_shim.reference_structure_loss()
    optim = torch.optim.AdamW([{"params": model.parameters(), "initial_lr": 1e-3}], lr=1e-3, weight_decay=5e-4)

    def train_step(autocast):
        model.train()
        optim.zero_grad()                                          # train.py:74-83
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            p0, p1, p2 = model(x)
        loss = loss_fn(p0.float(), mask) + loss_fn(p1.float(), mask) + loss_fn(p2.float(), mask)
        loss.item()
        loss.backward()
        optim.step()

    def infer_step(autocast):
        model.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            model(x)

    for mode, fn in (("train", train_step), ("infer", infer_step)):
        for autocast in (False, True):
            for _ in range(3):
                fn(autocast)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                fn(autocast)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / args.steps
            print(json.dumps({"what": "unmodified reference modules, PyTorch eager on cuda:0", "mode": mode,
                              "dtype": "autocast bf16" if autocast else "fp32 (TF32 off)", "variant": args.variant,
                              "batch": args.batch, "size": args.size, "ms_per_step": ms,
                              "img_per_s": args.batch / ms * 1e3, "torch": torch.__version__,
                              "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}), flush=True)


if __name__ == "__main__":
    main()
