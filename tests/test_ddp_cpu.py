"""world_size-2 gloo tests (CPU) of the data-parallel host logic: gradient buckets are reduced in the order the
engine reports them and cover exactly the active range; batch sharding; bench.py's reference arm under torchrun."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from sam2_unet_b200 import SAM2UNet
    from sam2_unet_b200.ddp import GradSync, shard_batch
    from sam2_unet_b200.model import FlatParams
    torch.manual_seed(0)
    m = SAM2UNet(model_cfg="tiny_test.yaml")
    flat = FlatParams(m, torch.device("cpu"))
    g = torch.Generator().manual_seed(100 + rank)
    flat.grad.copy_(torch.randn(flat.n_total, generator=g))
    mine = flat.grad.clone()
    sync = GradSync(flat.grad)
    nblocks = len(m.cfg.blocks)
    # replay the engine's notification order: decoder bucket first, then buckets as block indices count down
    for lo, hi, ready in flat.buckets:
        if ready == nblocks:
            sync.on_bucket(lo, hi)
    for i in range(nblocks - 1, -1, -1):
        for lo, hi, ready in flat.buckets:
            if ready == i:
                sync.on_bucket(lo, hi)
    covered = sorted(sync.ranges)
    sync.finish()
    other = torch.randn(flat.n_total, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    ok_sum = torch.allclose(flat.grad[:flat.n_active], (mine + other)[:flat.n_active], atol=1e-6)
    ok_tail = torch.equal(flat.grad[flat.n_active:], mine[flat.n_active:])          # up4.*: never reduced
    ok_cover = covered[0][0] == 0 and covered[-1][1] == flat.n_active and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    x = torch.arange(8).view(8, 1)
    ok_shard = shard_batch(x).flatten().tolist() == list(range(rank * 4, rank * 4 + 4))
    if rank == 0:
        json.dump(dict(sum=ok_sum, tail=ok_tail, cover=ok_cover, shard=ok_shard), open(out, "w"))
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks(tmp_path):
    out = str(tmp_path / "r.json")
    mp.spawn(_worker, args=(2, 29731, out), nprocs=2, join=True)
    res = json.load(open(out))
    assert all(res.values()), res


def test_bench_reference_arm_prints_one_line_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""                      # non-zero ranks exit 0 without work
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--cfg", "tiny_test.yaml", "--size", "96", "--cpu-batch", "2"], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "img/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
