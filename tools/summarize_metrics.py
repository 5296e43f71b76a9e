"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,
lts__t_bytes.sum,sm__pipe_tensor_cycles_active...` CSV of one optimisation step: per-kernel table and
the average DRAM traffic per tensor-core conv launch (what bench.py reports as roofline.traffic).

    python tools/summarize_metrics.py profiles/r1_ncu_conv_gram_metrics_512_v2.csv [--json KEY]
"""
from __future__ import annotations

import csv
import json
import sys
from collections import defaultdict
from pathlib import Path


def load(path: str) -> list[dict]:
    rows: dict[str, dict] = {}
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for rec in csv.DictReader(lines):
        r = rows.setdefault(rec["ID"], {"kernel": rec["Kernel Name"].split("(")[0], "grid": rec["Grid Size"]})
        r[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
    return [rows[k] for k in sorted(rows, key=int)]


def main() -> None:
    path = sys.argv[1]
    rows = load(path)
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    conv_bytes, conv_n, t_w, t_sum = 0.0, 0, 0.0, 0.0
    for r in rows:
        name = r["kernel"].replace("void stv::", "").replace("stv::", "")
        dram = r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)
        dur = r.get("gpu__time_duration.sum", 0)
        tens = r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)
        a = agg[name]
        a[0] += 1; a[1] += dur; a[2] += dram; a[3] += tens * dur
        if "conv_igemm" in name:
            conv_bytes += dram; conv_n += 1; t_w += tens * dur; t_sum += dur
    print(f"{'kernel':58s} {'n':>3s} {'us':>9s} {'DRAM MB':>9s} {'GB/s':>8s} {'tensor %':>8s}")
    for name, (n, dur, dram, tw) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:58]:58s} {n:3d} {dur/1e3:9.1f} {dram/1e6:9.1f} {dram/max(dur,1):8.0f} {tw/max(dur,1):8.1f}")
    if conv_n:
        avg = conv_bytes / conv_n
        print(f"conv launches: {conv_n}, average DRAM traffic per launch {avg/1e6:.1f} MB, "
              f"time-weighted tensor-pipe active {t_w/t_sum:.1f} %")
        if "--json" in sys.argv:
            key = sys.argv[sys.argv.index("--json") + 1]
            jf = Path(__file__).resolve().parent.parent / "profiles" / "roofline_traffic.json"
            data = json.loads(jf.read_text()) if jf.exists() else {}
            data[key] = avg
            jf.write_text(json.dumps(data))
            print(f"wrote {key} -> {jf}")


if __name__ == "__main__":
    main()
