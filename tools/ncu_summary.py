#!/usr/bin/env python3
"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into tracked files under profiles/.

    python tools/ncu_summary.py --round r1 gpurun_out/prof_verify.ncu-rep gpurun_out/prof_kv.ncu-rep ...
    python tools/ncu_summary.py --round r1 --launches gpurun_out/launches.csv

Writes profiles/<round>_<report>.md (one table per profiled launch) and merges the per-launch DRAM
traffic of each kernel into profiles/ncu_traffic.json, which bench.py reports as `roofline.traffic`.
Runs here (no GPU): it only reads reports with `ncu -i`.
"""
import argparse
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PROFILES = ROOT / "profiles"

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves / SM"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]

UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short_name(kernel: str) -> str:
    m = re.search(r"(\w+)<", kernel) or re.search(r"(\w+)\(", kernel)
    return m.group(1) if m else kernel


def read_raw(rep: Path):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        raise SystemExit(f"ncu -i {rep} failed: {out.stderr[:400]}")
    rows = list(csv.reader(io.StringIO(out.stdout)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def summarise_report(rep: Path, round_name: str, traffic: dict) -> Path:
    hdr, units, rows = read_raw(rep)
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {rep.name} ({round_name})", "",
             "`ncu --set full --clock-control none --import-source on`, one table per profiled launch. "
             "Durations under ncu are cold-cache and serialised: use them for shares and counters, not as bench values.", ""]
    per_kernel = defaultdict(list)
    for r in rows:
        kname = r[col["Kernel Name"]]
        lines += [f"## `{kname}`", "", "| metric | value | unit |", "|---|---|---|"]
        for key, label in METRICS:
            if key in col:
                lines.append(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |")
        if "dram__bytes_read.sum" in col:
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT_SCALE.get(units[col["dram__bytes_read.sum"]], 1)
            wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT_SCALE.get(units[col["dram__bytes_write.sum"]], 1)
            dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
            du = units[col["gpu__time_duration.sum"]]
            dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(du, 1.0)
            lines.append(f"| **DRAM traffic (read + write)** | {rd + wr:.0f} | byte |")
            lines.append(f"| **DRAM GB/s under ncu** | {(rd + wr) / dur_us / 1e3:.0f} | GB/s |")
            per_kernel[short_name(kname)].append({"dram_bytes": rd + wr, "duration_us": dur_us, "kernel": kname})
        lines.append("")
    for k, v in per_kernel.items():
        best = min(v, key=lambda d: d["duration_us"])
        traffic[k] = {"dram_bytes": int(sum(d["dram_bytes"] for d in v) / len(v)), "launches": len(v),
                      "duration_us_under_ncu": best["duration_us"], "report": rep.name, "round": round_name, "kernel": best["kernel"]}
    out = PROFILES / f"{round_name}_{rep.stem}.md"
    out.write_text("\n".join(lines))
    return out


def summarise_launches(csv_path: Path, round_name: str) -> Path:
    text = csv_path.read_text()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = defaultdict(list)
    order = []
    other = "(torch setup kernels outside the timed region: fill / randn / copy / index)"
    for r in rows:
        k = r["Kernel Name"]
        if "sd::" not in k:
            r["Kernel Name"] = k = other
            r["Grid Size"], r["Block Size"] = "-", "-"
        if k not in agg:
            order.append(k)
        agg[k].append(float(r["Metric Value"]) / 1e3)
    total = sum(sum(v) for k, v in agg.items() if k != other)
    lines = [f"# launch list {csv_path.name} ({round_name})", "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py`; per-launch times are cold-cache and "
             "serialised, so the SHARE column is the quantity to compare with the bench's event timings.", "",
             "| kernel | launches | grid x block | mean us | min us | max us | share of this library's GPU time |", "|---|---|---|---|---|---|---|"]
    geo = {r["Kernel Name"]: f'{r["Grid Size"]} x {r["Block Size"]}' for r in rows}
    for k in order:
        v = agg[k]
        share = "-" if k == other else f"{100 * sum(v) / total:.1f} %"
        lines.append(f"| `{k}` | {len(v)} | {geo[k]} | {sum(v) / len(v):.2f} | {min(v):.2f} | {max(v):.2f} | {share} |")
    out = PROFILES / f"{round_name}_{csv_path.stem}.md"
    out.write_text("\n".join(lines) + "\n")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r1")
    ap.add_argument("--launches", default=None)
    ap.add_argument("reports", nargs="*")
    a = ap.parse_args()
    PROFILES.mkdir(exist_ok=True)
    tpath = PROFILES / "ncu_traffic.json"
    traffic = json.loads(tpath.read_text()) if tpath.exists() else {}
    for rep in a.reports:
        print(summarise_report(Path(rep), a.round, traffic))
    if a.reports:
        tpath.write_text(json.dumps(traffic, indent=1, sort_keys=True) + "\n")
    if a.launches:
        print(summarise_launches(Path(a.launches), a.round))
    return 0


if __name__ == "__main__":
    sys.exit(main())
