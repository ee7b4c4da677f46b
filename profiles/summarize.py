"""Turn ncu outputs into the markdown summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv STEPS > profiles/rNN_launches.md
      launches.csv = `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python bench.py ...`
      STEPS        = number of training steps the profiled command ran (warm-up + timed + e2e + per-phase profile step)
  python profiles/summarize.py full gpurun_out/full_raw.csv > profiles/rNN_ncu_full.md
      full_raw.csv = `ncu -i capture.ncu-rep --page raw --csv`
  python profiles/summarize.py json gpurun_out/full_raw.csv profiles/rNN_ncu_full.json
      per-launch dram bytes / duration / tensor-pipe numbers as JSON: bench.py reads `roofline.traffic` from this file
"""
import csv
import io
import re
import sys
from collections import OrderedDict


def read_ncu_csv(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    if start < 0:
        raise SystemExit(f"{path}: no ncu CSV header found")
    return list(csv.DictReader(io.StringIO(text[start:])))


def short(name):
    name = re.sub(r"\(.*$", "", name)          # drop the argument list
    name = name.replace("void ", "")
    name = re.sub(r"cub::CUB_\d+_NS::", "cub::", name)
    name = re.sub(r"cub::detail::\w+::", "cub::", name)
    return name[:90]


def launches(path, steps):
    rows = [r for r in read_ncu_csv(path) if r.get("Metric Name") == "gpu__time_duration.sum"]
    if steps == "auto":
        # steps only: drop the set-up kernels before the first step's first kernel (pack_dense) and the trailing partial step
        # (after the last dense-optimiser launch); one sigmoid_bce launch per step
        first = next((i for i, r in enumerate(rows) if "pack_dense_kernel" in r["Kernel Name"]), 0)
        last = max((i for i, r in enumerate(rows) if "adam_kernel" in r["Kernel Name"] or "sgd_kernel" in r["Kernel Name"]), default=len(rows) - 1)
        rows = rows[first : last + 1]
        steps = max(1, sum(1 for r in rows if "sigmoid_bce_kernel" in r["Kernel Name"]))
        print(f"{steps} training steps captured (set-up kernels before the first step and the trailing partial step dropped)\n")
    steps = float(steps)
    agg = OrderedDict()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit.startswith("n") else (v if unit.startswith("u") else v * 1e3)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches/step | avg µs | µs/step | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n / steps:.1f} | {us / n:.1f} | {us / steps:.1f} | {100 * us / total:.1f}% |")
    print(f"| **total** | {sum(a[0] for a in agg.values()) / steps:.1f} |  | {total / steps:.1f} | 100% |")


FULL_METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor-pipe active %"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu.sum", "smem bank conflicts"),
    ("smsp__cycles_active.avg", "SM active cycles"),
]


def full(path):
    rows = read_ncu_csv(path)
    # --page raw --csv: one row per launch, one column per metric (second line holds the units)
    units = rows[0]
    cols = [c for c in rows[0].keys()]
    print("| kernel | " + " | ".join(label for _, label in FULL_METRICS) + " |")
    print("|---|" + "---:|" * len(FULL_METRICS))
    for r in rows[1:]:
        cells = []
        for m, _ in FULL_METRICS:
            c = next((c for c in cols if c == m or c.startswith(m)), None)
            if c is None or r.get(c, "") == "":
                cells.append("–")
                continue
            cells.append(f"{r[c]} {units.get(c, '')}".strip())
        print(f"| `{short(r.get('Kernel Name', '?'))}` | " + " | ".join(cells) + " |")


def _num(cell, unit):
    v = float(str(cell).replace(",", ""))
    u = (unit or "").lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
             "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    return v * scale.get(u, 1.0)


def to_json(path, out):
    import json

    rows = read_ncu_csv(path)
    units, cols = rows[0], list(rows[0].keys())

    def col(m):
        return next((c for c in cols if c == m or c.startswith(m)), None)

    ker = []
    for r in rows[1:]:
        try:
            ker.append({"name": short(r.get("Kernel Name", "?")),
                        "duration_us": _num(r[col("gpu__time_duration.sum")], units[col("gpu__time_duration.sum")]),
                        "dram_read_bytes": _num(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")]),
                        "dram_write_bytes": _num(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")]),
                        "tensor_pipe_active_pct": float(r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")] or 0),
                        "warps_active_pct": float(r[col("sm__warps_active.avg.pct_of_peak_sustained_active")] or 0),
                        "regs_per_thread": float(r[col("launch__registers_per_thread")] or 0)})
        except (KeyError, TypeError, ValueError):
            continue
    json.dump({"source": path, "kernels": ker}, open(out, "w"), indent=1)
    print(f"{out}: {len(ker)} launches")


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "json":
        to_json(sys.argv[2], sys.argv[3])
        raise SystemExit(0)
    if len(sys.argv) >= 4 and sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "auto")
    elif len(sys.argv) >= 3 and sys.argv[1] == "full":
        full(sys.argv[2])
    else:
        raise SystemExit(__doc__)
