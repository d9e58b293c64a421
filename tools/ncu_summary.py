"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --csv` launch list (one eager step of bench.py) per kernel:
launches, total / share of device time, average DRAM bytes per launch (read + write), time-weighted tensor-pipe activity.

usage: python tools/ncu_summary.py <launches.csv> <key, e.g. tf32_b32> [profiles/ncu_kernel_summary_r2.json] [shares.txt]
The JSON (merged under <key>) is what bench.py reads to fill roofline.traffic / tensor_pipe_pct."""
import csv
import json
import os
import re
import sys


def parse(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    per = {}
    for r in rd:
        kid = r["ID"]
        d = per.setdefault(kid, {"name": r["Kernel Name"]})
        val = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            d["ms"] = val * scale
        elif m.startswith("dram__bytes"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d[m] = val * scale
        elif m.startswith("sm__pipe_tensor"):
            d["tensor_pct"] = val
    return list(per.values())


def main():
    path, key = sys.argv[1], sys.argv[2]
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    out_txt = sys.argv[4] if len(sys.argv) > 4 else None
    rows = parse(path)
    total = sum(r.get("ms", 0.0) for r in rows)
    agg = {}
    for r in rows:
        name = re.sub(r"<.*", "", r["name"].split("(")[0]).split("::")[-1].strip() or r["name"][:40]
        a = agg.setdefault(name, {"launches": 0, "ms": 0.0, "dram": 0.0, "tw": 0.0})
        a["launches"] += 1
        a["ms"] += r.get("ms", 0.0)
        a["dram"] += r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
        a["tw"] += r.get("tensor_pct", 0.0) * r.get("ms", 0.0)
    summ = {n: {"launches": a["launches"], "total_ms": round(a["ms"], 4), "share_of_step": round(a["ms"] / total, 4),
                "dram_bytes_per_launch": round(a["dram"] / a["launches"], 1),
                "dram_gb_per_s": round(a["dram"] / max(a["ms"], 1e-9) / 1e6, 1),
                "tensor_pipe_pct": round(a["tw"] / max(a["ms"], 1e-9), 2)} for n, a in agg.items()}
    lines = [f"{len(rows)} launches, {total:.2f} ms (ncu: cold-cache, serialised; compare shares)  [{key}]"]
    for n, v in sorted(summ.items(), key=lambda kv: -kv[1]["total_ms"]):
        lines.append(f"{v['total_ms']:9.3f} ms {100 * v['share_of_step']:5.1f}% n={v['launches']:4d}  dram/launch {v['dram_bytes_per_launch'] / 1e6:9.2f} MB"
                     f"  {v['dram_gb_per_s']:7.0f} GB/s  tensor {v['tensor_pipe_pct']:5.1f}%  {n}")
    print("\n".join(lines))
    if out_txt:
        open(out_txt, "w").write("\n".join(lines) + "\n")
    if out_json:
        d = json.load(open(out_json)) if os.path.exists(out_json) else {}
        d[key] = summ
        json.dump(d, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
