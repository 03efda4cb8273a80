#!/usr/bin/env python
"""bench.py — SAM2-UNet Hiera-L 352x352 train-step (and inference) throughput on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|infer] [--impl b200|reference]

Metric (BASELINE.json): images/s of the full train step (forward + 3 x structure_loss + backward + AdamW),
SAM2-UNet Hiera-L, 352x352, batch 12 per GPU, bf16 compute, synthetic data, random-init weights.
N > 1 is launched by torchrun (one rank per GPU, NCCL); the batch is sharded (weak scaling: 12 images per GPU)
and the only collective is the bucketed gradient all-reduce overlapped with backward.

One JSON line on rank 0:
  value         device-resident throughput (inputs already in HBM), whole job, CUDA-event timed, max over ranks
  e2e           same metric through the public TrainStep call with HOST (pinned) batches: H2D of the batch and a
                D2H read of the loss inside the timed region, every step
  infer         the other half of BASELINE.json's metric in the same run: inference forward img/s (device-resident and
                end to end from pinned host batches) of the same model at the same batch
  roofline      the dominant tcgen05 GEMM shape: algorithmic FLOPs / CUDA-event duration, both ways the recipe names -
                launched alone back to back against the BURST bf16 peak (`achieved`, `frac`), and inside the step
                against the SUSTAINED peak (`achieved_in_step`, `frac_in_step`)
  cpu_baseline  N = 1 only: the reference's own modules (staged under baseline/_ref by scripts/stage_reference.py,
                imported through oracle/ref_shim.py; kind "reference") - or the oracle port when they are absent (kind
                "port") - timed on this box's host cores on a bounded sample (one 12-image step) of the same workload
`--impl reference` times that CPU path alone and prints the same line shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# (M, N, K) -> DRAM bytes of one launch measured by ncu --set full (see profiles/); filled per round
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant GEMM (stage-3 fc1, GELU + saved GELU'
# epilogue) from profiles/r1_gemm_pair_full_raw.csv rows 3 and 9: 9.41 MB read + 3.2 / 2.6 MB written before the kernel
# ends (the 126 MB L2 absorbs the rest of the 53 MB of output)
NCU_TRAFFIC = {(5808, 2304, 576): 12.3e6}

METRIC = {"train": "train img/s, SAM2-UNet Hiera-L 352x352 (fwd + 3x structure_loss + bwd + AdamW)",
          "infer": "infer img/s, SAM2-UNet Hiera-L 352x352 forward"}


def metric_name(args):
    """BASELINE.json's metric on the default configuration; other trunks / sizes are named as what they are."""
    if args.cfg == "sam2_hiera_l.yaml" and args.size == 352:
        return METRIC[args.mode]
    trunk = args.cfg.replace("sam2_hiera_", "Hiera-").replace(".yaml", "").replace("Hiera-l", "Hiera-L")
    return METRIC[args.mode].replace("Hiera-L 352x352", f"{trunk} {args.size}x{args.size}")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--batch", type=int, default=12, help="images per GPU")
    ap.add_argument("--size", type=int, default=352)
    ap.add_argument("--cfg", default="sam2_hiera_l.yaml")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=0, help="images per CPU-baseline step (0 = --batch: the same config)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: total images per step, split evenly over the ranks (0 = weak: --batch per GPU)")
    ap.add_argument("--no-infer", action="store_true", help="skip the inference half of the line")
    ap.add_argument("--profile-out", default="", help="write the per-op CUDA-event breakdown of one eager step here")
    ap.add_argument("--ncu-range", action="store_true",
                    help="after warm-up run ONE step between cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="B200_PROFILING.md fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU reference arm

def _reference_root():
    """Where the unmodified reference can be imported from: the build container's checkout or the staged copy."""
    for root in (os.environ.get("SAM2UNET_REFERENCE", ""), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if root and os.path.isfile(os.path.join(root, "SAM2UNet.py")):
            return root
    return None


_VARIANT = {"sam2_hiera_l.yaml": "l", "sam2_hiera_s.yaml": "s", "sam2_hiera_t.yaml": "t", "sam2_hiera_b+.yaml": "b+",
            "tiny_test.yaml": "test"}


def cpu_step_fn(cfg_key: str, batch: int, size: int, mode: str):
    """-> (step, kind).  One step of the reference on the host cores: its own SAM2UNet / Hiera modules, its
    structure_loss and torch.optim.AdamW exactly as train.py:48-52,74-83 drives them (kind "reference"); when the
    reference files are not on this box, the oracle port of the same path (kind "port")."""
    import torch
    from sam2_unet_b200.params import fill_deterministic_
    from sam2_unet_b200.synthetic import synthetic_batch
    torch.set_num_threads(os.cpu_count() or 1)
    x, mask = synthetic_batch(batch, size, seed=0)
    root = _reference_root()
    if root is not None and cfg_key != "tiny_test.yaml":
        os.environ["SAM2UNET_REFERENCE"] = root
        from oracle import ref_shim
        model = ref_shim.build_reference(_VARIANT[cfg_key])
        fill_deterministic_(model, 0)
        loss_fn = ref_shim.reference_structure_loss()
        optim = torch.optim.AdamW([{"params": model.parameters(), "initial_lr": 1e-3}], lr=1e-3, weight_decay=5e-4)

        def step():
            if mode == "infer":
                model.eval()
                with torch.no_grad():
                    model(x)
                return
            model.train()
            optim.zero_grad()                              # train.py:74-83
            pred0, pred1, pred2 = model(x)
            loss = loss_fn(pred0, mask) + loss_fn(pred1, mask) + loss_fn(pred2, mask)
            loss.item()
            loss.backward()
            optim.step()
        return step, "reference"

    from oracle import port
    from sam2_unet_b200 import SAM2UNet
    m = SAM2UNet(model_cfg=cfg_key)                     # parameter container only (never run on the CPU)
    fill_deterministic_(m, 0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    del m
    trunk = port.TRUNKS[_VARIANT[cfg_key]]
    keys = port.trainable_keys(sd)
    mom = {k: torch.zeros_like(sd[k]) for k in keys}
    var = {k: torch.zeros_like(sd[k]) for k in keys}
    state = {"t": 0}

    def step():
        if mode == "infer":
            with torch.no_grad():
                port.forward(sd, trunk, x, False)
            return
        state["t"] += 1
        bn = port.BNState()
        _, _, grads = port.loss_and_grads(sd, trunk, x, mask, True, bn)
        for k in keys:
            if grads[k] is not None:
                port.adamw_step(sd[k], grads[k], mom[k], var[k], state["t"])
        sd.update(bn.updates)

    return step, "port"


def time_cpu(args, steps: int, warmup: int, budget_s: float):
    from sam2_unet_b200.config import canonical_name
    batch = args.cpu_batch or args.batch
    step, kind = cpu_step_fn(canonical_name(args.cfg), batch, args.size, args.mode)
    t_begin = time.perf_counter()
    for _ in range(max(1, warmup)):
        step()
        if time.perf_counter() - t_begin > budget_s / 3:  # bounded: one warm-up step is enough on a slow host
            break
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    total = sum(times)
    return batch * len(times) / total, total / len(times), len(times), kind, batch


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, sec, done, kind, batch = time_cpu(args, args.steps, min(args.warmup, 1), 150.0)
    cores = os.cpu_count() or 1
    what = "the reference's own modules (SAM2UNet.py, hieradet.py, train.py:structure_loss, torch.optim.AdamW)" \
        if kind == "reference" else "oracle port of the reference"
    sample = (f"{done} timed step(s) of {batch} images after 1 warm-up (bounded to 150 s of the requested {args.steps}), "
              f"{what} on torch-CPU fp32, {cores} threads")
    line = {"impl": "reference", "metric": metric_name(args), "value": ips, "unit": "img/s", "n_gpus": args.gpus,
            "steps": done, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "cpu_batch": batch},
            "cpu_baseline": {"value": ips, "unit": "img/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"SAM2-UNet {args.cfg} {args.size}x{args.size} {'train step (fwd+bwd+AdamW, structure_loss)' if args.mode == 'train' else 'forward'}, "
            f"batch {args.batch}/GPU, {args.dtype}")


# --------------------------------------------------------------------------------------------------- B200 arm

def gemm_flops(rec):
    """Algorithmic FLOPs and time of the GEMM-family calls of one instrumented step, and the same for the single
    (M, N, K) shape that takes the most time (the stage-3 MLP GEMM of Hiera-L)."""
    fl = ms = 0.0
    n = 0
    shapes = {}
    for name, t, a in rec:
        if name == "s2u_gemm":
            fl += 2.0 * a[6] * a[7] * a[8]
            ms += t
            n += 1
            s = shapes.setdefault((a[6], a[7], a[8]), [0, 0.0])
            s[0] += 1
            s[1] += t
    top = max(shapes.items(), key=lambda kv: kv[1][1]) if shapes else None
    return fl, ms, n, top


def gemm_variants(rec, shape):
    """Distinct epilogue variants (flags + which optional operands are present) of the s2u_gemm calls of `shape`
    in one instrumented step, with their call counts."""
    out = {}
    for name, _, a in rec:
        if name == "s2u_gemm" and (a[6], a[7], a[8]) == shape:
            key = (a[16], bool(a[9]), bool(a[10]), bool(a[12]), bool(a[14]), a[17], a[18])
            out[key] = out.get(key, 0) + 1
    return out


def gemm_back_to_back(dev, shape, variants, iters=40, sets=4):
    """Average duration of the dominant GEMM launched back to back, CUDA events on the launching stream, every
    epilogue variant the step uses weighted by its call count.  The operands rotate through `sets` buffer sets
    (> 126 MB in total) so no launch finds its inputs in L2 from the previous one."""
    import torch
    from sam2_unet_b200 import _lib
    M, N, K = shape
    st = torch.cuda.current_stream(dev).cuda_stream
    tot_ms = tot_calls = 0.0
    detail = []
    for (flags, has_bias, has_pre, has_aux, has_res, dt, backend), calls in sorted(variants.items()):
        el = torch.bfloat16 if dt == 1 else torch.float32
        bufs = []
        for _ in range(sets):
            A = torch.randn(M, K, device=dev).to(el)
            W = (torch.randn(N, K, device=dev) * 0.05).to(el)
            C = torch.empty(M, N, device=dev, dtype=torch.float32 if flags & 16 else el)
            bias = torch.randn(N, device=dev) if has_bias else None
            pre = torch.empty(M, N, device=dev, dtype=el) if has_pre else None
            aux = torch.randn(M, N, device=dev).to(el) if has_aux else None
            res = torch.randn(M, N, device=dev).to(torch.float32 if flags & 32 else el) if has_res else None
            bufs.append((A, W, C, bias, pre, aux, res))

        def launch(b):
            A, W, C, bias, pre, aux, res = b
            p = lambda t: t.data_ptr() if t is not None else 0          # noqa: E731
            _lib.call("s2u_gemm", A.data_ptr(), K, W.data_ptr(), K, C.data_ptr(), N, M, N, K, p(bias), p(pre),
                      N if pre is not None else 0, p(aux), N if aux is not None else 0, p(res),
                      N if res is not None else 0, flags, dt, backend, st)
        for i in range(8):
            launch(bufs[i % sets])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for i in range(iters):
            launch(bufs[i % sets])
        e1.record()
        torch.cuda.synchronize(dev)
        us = e0.elapsed_time(e1) / iters * 1e3
        detail.append({"flags": flags, "calls_per_step": calls, "us": us})
        tot_ms += us * calls
        tot_calls += calls
        del bufs
    return tot_ms / tot_calls, detail


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    from sam2_unet_b200 import Predictor, SAM2UNet, TrainStep, _lib
    from sam2_unet_b200.params import fill_deterministic_
    from sam2_unet_b200.synthetic import synthetic_batch

    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        args.batch = args.global_batch // world
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:       # N = 1 only (torchrun pins OMP_NUM_THREADS=1)
        ips, sec, done, kind, batch = time_cpu(args, 1, 1, 60.0)
        cores = os.cpu_count() or 1
        cpu = {"value": ips, "unit": "img/s", "cores": cores, "kind": kind,
               "sample": f"{done} timed step of {batch} images after 1 warm-up, "
                         f"{'the reference modules' if kind == 'reference' else 'oracle port'} (torch-CPU fp32), {cores} threads"}

    torch.manual_seed(0)
    model = SAM2UNet(model_cfg=args.cfg, dtype=args.dtype)
    fill_deterministic_(model, 0)
    model = model.to(dev)
    B, S = args.batch, args.size
    xh, mh = synthetic_batch(B, S, seed=100 + rank)
    xh, mh = xh.pin_memory(), mh.pin_memory()
    xd, md = xh.to(dev), mh.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if args.mode == "train":
        model.train()
        step = TrainStep(model, lr=1e-3, weight_decay=5e-4, use_graph=not args.no_graph)
        run_dev = lambda: step(xd, md)                          # noqa: E731
        def run_e2e():                                          # H2D batch + D2H loss every step
            loss = step(xh, mh)                                 # picks up the copy prefetched during the previous step
            step.prefetch(xh, mh)                               # next batch's H2D overlaps this step's kernels
            return loss.cpu()
        d2h = 12
    else:
        model.eval()
        pred = Predictor(model, use_graph=not args.no_graph)
        run_dev = lambda: pred(xd)                              # noqa: E731
        def run_e2e():                                          # H2D batch + D2H of the main logits map
            out = pred(xh)[0]
            pred.prefetch(xh)
            return out.cpu()
        d2h = B * S * S * 4
    h2d = xh.numel() * 4 + (mh.numel() * 4 if args.mode == "train" else 0)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, clk.summary(), _lib.launch_count() - l0

    warm = max(args.warmup, 3)
    if args.ncu_range:
        for _ in range(warm + 3):
            run_dev()
        torch.cuda.synchronize(dev)
        torch.cuda.cudart().cudaProfilerStart()
        run_dev()
        torch.cuda.synchronize(dev)
        torch.cuda.cudart().cudaProfilerStop()
        print("ncu range done", flush=True)
        return
    ms, clocks, _ = timed(run_dev, args.steps, warm + 3)        # +3: two eager warm-ups and the graph capture
    value = B * world * args.steps / (ms / 1e3)
    ms_e2e, _, _ = timed(run_e2e, args.steps, 2)
    e2e = B * world * args.steps / (ms_e2e / 1e3)

    # the other half of the metric: inference forward of the same model at the same batch (graph-replayed Predictor)
    infer = None
    if args.mode == "train" and not args.no_infer:
        model.eval()
        pred = Predictor(model, use_graph=not args.no_graph)

        def infer_e2e():
            out = pred(xh)[0]
            pred.prefetch(xh)
            return out.cpu()
        ms_i, _, _ = timed(lambda: pred(xd), args.steps, warm + 3)
        ms_ie, _, _ = timed(infer_e2e, args.steps, 2)
        infer = {"metric": METRIC["infer"] if metric_name(args) == METRIC["train"] else metric_name(args).replace("train", "infer"),
                 "value": B * world * args.steps / (ms_i / 1e3), "unit": "img/s", "ms_per_step": ms_i / args.steps,
                 "e2e": {"value": B * world * args.steps / (ms_ie / 1e3), "unit": "img/s",
                         "h2d_bytes_per_step": xh.numel() * 4, "d2h_bytes_per_step": B * S * S * 4,
                         "ms_per_step": ms_ie / args.steps}}
        del pred
        model.train()

    # instrumented eager step: per-op CUDA-event durations (kernel shares, roofline of the GEMM family)
    roof = None
    launches_per_step = None
    if rank == 0:
        pk = peaks()
        if args.mode == "train":
            eager = TrainStep(model, lr=1e-3, weight_decay=5e-4, use_graph=False, sync_grads=False)
            eager.optim = step.optim
            fn = lambda: eager(xd, md)                          # noqa: E731
        else:
            def fn():
                with torch.no_grad():
                    return model(xd)
        # the instrumented step runs with the decoder's side streams off, so that every duration is that kernel alone
        # on the GPU (the ncu launch list under profiles/ is serialised the same way)
        eng = model._engine(dev)
        was_overlap, eng.overlap = eng.overlap, False
        fn()
        torch.cuda.synchronize(dev)
        _lib.profile_begin()
        fn()
        rec = _lib.profile_end()
        eng.overlap = was_overlap
        launches_per_step = len(rec)
        fl, gms, ng, top = gemm_flops(rec)
        total_ms = sum(t for _, t, _ in rec)
        (tm, tn, tk), (tcalls, tms) = top
        tfl = 2.0 * tm * tn * tk
        # `achieved` uses the kernel launched back to back (every epilogue variant of the step, weighted by calls);
        # the per-launch event brackets of the eager step add the launch latency of an idle GPU to every call and are
        # reported next to it as us_per_launch_in_step
        b2b_us, b2b_detail = gemm_back_to_back(dev, (tm, tn, tk), gemm_variants(rec, (tm, tn, tk)))
        achieved = tfl / (b2b_us * 1e-6) / 1e12
        in_step = tfl * tcalls / (tms * 1e-3) / 1e12 if tms > 0 else 0.0
        all_tf = fl / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of one launch of this shape from the committed
        # `ncu --set full` capture (profiles/r1_gemm_pair_full_raw.csv); None when the shape differs from the captured one
        traffic = NCU_TRAFFIC.get((tm, tn, tk))
        roof = {"bound": "tensor", "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                "frac": achieved / pk["tf_burst"], "traffic": traffic,
                "frac_in_step": in_step / pk["tf_sustained"], "peak_in_step": pk["tf_sustained"],
                "kernel": f"gemm_umma_pair_kernel (persistent CTA-pair tcgen05 + TMA GEMM), dominant shape M={tm} N={tn} K={tk}",
                "algorithmic_flop_per_launch": tfl, "launches_of_shape_per_step": tcalls,
                "us_per_launch": b2b_us, "variants": b2b_detail,
                "us_per_launch_in_step": tms / tcalls * 1e3, "achieved_in_step": in_step,
                "all_gemm_launches": ng, "all_gemm_tflops": all_tf, "gemm_ms_per_step": gms,
                "gemm_share_of_step": gms / total_ms if total_ms else None,
                "timing": "achieved / frac: the kernel timed ALONE - CUDA events around 40 back-to-back launches per epilogue "
                          "variant on the launching stream, operands rotating through 4 buffer sets (> L2) - against the "
                          "BURST bf16 peak; achieved_in_step / frac_in_step: CUDA events around each launch of one eager "
                          "step (adds ~5 us idle-GPU launch latency per call) against the SUSTAINED peak",
                "peak_source": pk["source"] + ": burst figure for the isolated timing, sustained for the in-step one"}
        if args.profile_out:
            by = {}
            for name, t, _ in rec:
                c = by.setdefault(name, [0, 0.0])
                c[0] += 1
                c[1] += t
            gs = {}
            for name, t, a in rec:
                if name == "s2u_gemm":
                    c = gs.setdefault(f"{a[6]}x{a[7]}x{a[8]} flags={a[16]}", [0, 0.0])
                    c[0] += 1
                    c[1] += t
            with open(args.profile_out, "w") as f:
                json.dump({"total_ms": total_ms, "by_op": {k: {"calls": v[0], "ms": v[1]} for k, v in
                                                             sorted(by.items(), key=lambda kv: -kv[1][1])},
                           "gemm_shapes": {k: {"calls": v[0], "ms": v[1], "us_per_call": v[1] / v[0] * 1e3}
                                           for k, v in sorted(gs.items(), key=lambda kv: -kv[1][1])}}, f, indent=1)
    if rank == 0:
        line = {"metric": metric_name(args), "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.global_batch else "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": workload_name(args), "global_batch": B * world, "parallelism": f"dp{world}",
                           "cuda_graph": not args.no_graph,
                           "untimed_steps_before_the_timed_region": warm + 3,
                           "l2": "per-step working set (activations + im2col buffers, several GB) is far larger than "
                                 "the 126 MB L2, no explicit flush"},
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": (launches_per_step or 0) * args.steps,
                "infer": infer, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        # leave without tearing NCCL down: destroying a communicator whose collectives were captured in a live CUDA
        # graph can block; everything has been reported and synchronised at this point
        torch.cuda.synchronize(dev)
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
