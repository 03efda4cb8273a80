"""Builds profiles/r1_summary.md from the committed ncu exports (launch list + full capture of the GEMM launches)."""
import collections
import csv
import io
import re
import sys

root = sys.argv[1] if len(sys.argv) > 1 else "profiles"
lines = [l for l in open(f"{root}/r1_launches.csv") if not l.startswith("==")]
rows = [r for r in csv.DictReader(io.StringIO("".join(lines))) if r.get("Metric Name") == "gpu__time_duration.sum"]
fam = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    n = re.sub(r"<.*", "", n)
    t = float(r["Metric Value"].replace(",", "")) / 1e3
    fam[n][0] += 1
    fam[n][1] += t
tot = sum(v[1] for v in fam.values())
out = []
out.append(f"| kernel family | launches | total us | share |\n|---|---:|---:|---:|")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
table1 = "\n".join(out)
groups = {"tcgen05 GEMM (`gemm_umma_pair_kernel`, `gemm_umma_ws_kernel`)": ("gemm_umma",),
          "tcgen05 weight-gradient GEMM": ("wgrad_umma",),
          "mma.sync attention": ("amma::",),
          "LayerNorm": ("ln_",), "BatchNorm": ("bn_",), "im2col / stem": ("im2col", "patch")}
gl = []
for name, keys in groups.items():
    t = sum(v[1] for k, v in fam.items() if any(x in k for x in keys))
    gl.append(f"{name} {100 * t / tot:.1f} %")
rows2 = list(csv.reader(open(f"{root}/r1_gemm_pair_full_raw.csv")))
hdr, body = rows2[0], rows2[2:]
c = hdr.index
cols = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "CTAs"), ("dram__bytes_read.sum", "DRAM rd MB"),
        ("dram__bytes_write.sum", "DRAM wr MB"), ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA ld MB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %")]
t2 = ["| # | instantiation | " + " | ".join(n for _, n in cols) + " |", "|---|---|" + "---:|" * len(cols)]
for i, r in enumerate(body):
    inst = re.search(r"gemm_umma_pair_kernel<[^>]*>", r[c("Kernel Name")]).group(0)
    t2.append(f"| {i} | `{inst}` | " + " | ".join(f"{float(r[c(h)]):.2f}" if h in hdr else "-" for h, _ in cols) + " |")
print(f"LAUNCHES={sum(v[0] for v in fam.values())} TOTAL_MS={tot / 1e3:.2f}")
print("GROUPS: " + "; ".join(gl))
open(f"{root}/_tables.md", "w").write(table1 + "\n\n" + "\n".join(t2) + "\n")
