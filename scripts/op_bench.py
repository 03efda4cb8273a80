"""GPU-box tool: warm-cache, launch-overhead-free duration of every distinct C-ABI call of one train step.
Records the (entry point, arguments) of one eager step, groups identical calls (same entry point and integer arguments),
and times ONE representative of each group as a CUDA-graph replay of 20 back-to-back launches on its original buffers
(L2-warm, like inside the step where the producer has just written them).  Prints calls x us per group."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam2_unet_b200.synthetic import synthetic_batch  # noqa: E402
from sam2_unet_b200 import SAM2UNet, TrainStep, _lib  # noqa: E402
from sam2_unet_b200.params import fill_deterministic_  # noqa: E402

# torch.cuda.graph() empties the caching allocator on entry, which would unmap the step's (freed) temporaries that the
# recorded calls point to
torch.cuda.empty_cache = lambda: None
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
model = SAM2UNet(model_cfg="sam2_hiera_l.yaml", dtype="bf16")
fill_deterministic_(model, 0)
model = model.to(dev).train()
x, m = synthetic_batch(B, 352, seed=0)
x, m = x.to(dev), m.to(dev)
step = TrainStep(model, lr=1e-3, weight_decay=5e-4, use_graph=False, sync_grads=False)
eng = model._engine(dev)
eng.overlap = False
step(x, m)
step(x, m)
torch.cuda.synchronize()
_lib.profile_begin()
step(x, m)
rec = _lib._profile
_lib._profile = None
torch.cuda.synchronize()
groups = collections.OrderedDict()
for name, _, _, args in rec:
    # pointers are large ints: keep only "small" integer / float arguments in the key
    key = (name,) + tuple(a for a in args if isinstance(a, float) or (isinstance(a, int) and abs(a) < (1 << 32)))
    g = groups.setdefault(key, [0, args])
    g[0] += 1
lib = _lib.load()
res = []
side = torch.cuda.Stream()
for key, (calls, args) in groups.items():
    name = key[0]
    fn = getattr(lib, name)
    st = side.cuda_stream
    a = list(args)
    a[-1] = st                                   # every entry point takes the stream last
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(20):
                    rc = fn(*a)
                    assert rc == 0, (name, rc)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
    except Exception as e:  # noqa: BLE001
        us = float("nan")
        print("skip", name, e)
    res.append((calls * us, calls, us, key))
tot = sum(r[0] for r in res if r[0] == r[0])
print(f"{len(rec)} calls, {len(groups)} distinct; sum of warm durations {tot / 1e3:.2f} ms")
by = collections.defaultdict(float)
for t, calls, us, key in res:
    by[key[0]] += t
for k, v in sorted(by.items(), key=lambda kv: -kv[1]):
    print(f"  {k:28s} {v / 1e3:7.3f} ms {100 * v / tot:5.1f}%")
print()
for t, calls, us, key in sorted(res, key=lambda r: -r[0])[:400]:
    print(f"{t / 1e3:7.3f} ms  {calls:4d} x {us:7.1f} us  {key[0]:24s} {key[1:]}")
