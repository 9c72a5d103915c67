#!/usr/bin/env python
"""Summarise an ncu launch list (csv with gpu__time_duration / dram__bytes_read / dram__bytes_write per launch) of
tools/prof_step.py: per kernel class launches, time share, DRAM bytes; optionally writes profiles/span_traffic.json
(the DRAM bytes of the custom-kernel span that bench.py's roofline.traffic reports).

    python tools/launch_summary.py gpurun_out/r2_launches.csv --steps 1 --skip-steps 1 [--write-span profiles/span_traffic.json]
"""
import argparse, collections, csv, json, re, sys

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--steps", type=int, default=1, help="timed steps in the capture")
ap.add_argument("--skip-steps", type=int, default=1, help="warm-up steps in the capture (dropped: first 1/(skip+steps) of each kernel's launches)")
ap.add_argument("--write-span", default=None)
a = ap.parse_args()
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3,
        "msecond": 1.0, "second": 1e3}
lines = [l for l in open(a.csv) if not l.startswith("==")]
per = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r"^void ", "", row["Kernel Name"])
    name = re.sub(r"\(.*", "", name)
    d = per.setdefault(name, collections.defaultdict(list))
    v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
    d[row["Metric Name"]].append(v)
tot_t = 0.0
rows = []
for name, d in per.items():
    n = len(d.get("gpu__time_duration.sum", [])) or len(d.get("dram__bytes_read.sum", []))
    keep = lambda xs: xs[len(xs) * a.skip_steps // (a.skip_steps + a.steps):] if xs else []
    t = sum(keep(d.get("gpu__time_duration.sum", []))) / a.steps
    rd = sum(keep(d.get("dram__bytes_read.sum", []))) / a.steps
    wr = sum(keep(d.get("dram__bytes_write.sum", []))) / a.steps
    rows.append((name, len(keep(d.get("gpu__time_duration.sum", d.get("dram__bytes_read.sum", [])))) / a.steps, t, rd, wr))
    tot_t += t
ours = lambda n: n.startswith(("hy::", "tc::", "pg::", "wg::", "ln::", "fx::")) or "hy::" in n
span = lambda n: ours(n) and not ("proj_" in n or "wgrad" in n or "ln::" in n)
print(f"{'kernel':78s} {'launches':>8s} {'ms':>8s} {'share':>6s} {'rd GB':>7s} {'wr GB':>7s}")
for name, n, t, rd, wr in sorted(rows, key=lambda r: -r[2]):
    print(f"{name[:78]:78s} {n:8.1f} {t:8.3f} {100 * t / max(tot_t, 1e-9):5.1f}% {rd / 1e9:7.2f} {wr / 1e9:7.2f}")
sp_r = sum(r[3] for r in rows if span(r[0])); sp_w = sum(r[4] for r in rows if span(r[0]))
sp_t = sum(r[2] for r in rows if span(r[0])); pj_t = sum(r[2] for r in rows if ours(r[0]) and not span(r[0]))
print(f"\nper step: all kernels {tot_t:.3f} ms (serialised, cold-ish caches); span kernels {sp_t:.3f} ms, DRAM {sp_r / 1e9:.2f} GB read + "
      f"{sp_w / 1e9:.2f} GB written = {(sp_r + sp_w) / 1e9:.2f} GB; projection / glue kernels of this library {pj_t:.3f} ms")
if a.write_span:
    json.dump({"what": "DRAM bytes (read+write) of the custom-kernel span (FFT passes + filter kernels) in one fwd+bwd step at L=2^20, D=256, "
                       "B=1, from ncu --cache-control none", "source": a.csv.replace("gpurun_out/", "profiles/"),
               "span_dram_bytes_per_step": sp_r + sp_w, "read": sp_r, "write": sp_w}, open(a.write_span, "w"))
    print("wrote", a.write_span)
