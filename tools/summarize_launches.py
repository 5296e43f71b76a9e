"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time shares."""
import collections
import csv
import sys


def summarize(path: str, per_launch: bool = False) -> str:
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
    tot, cnt, seq = {}, collections.Counter(), []
    for r in rows[1:]:
        k = r[ki].split("(")[0][:70]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v
        tot[k] = tot.get(k, 0) + v
        cnt[k] += 1
        seq.append((k, v, r[gi] if gi is not None else ""))
    s = sum(tot.values())
    out = [f"total {s:.1f} us over {sum(cnt.values())} launches"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append(f"{v:10.1f} us {100 * v / s:5.1f}%  x{cnt[k]:3d}  {k}")
    if per_launch:
        out.append("-- launch order --")
        out += [f"{v:9.1f} us  grid {g:>14s}  {k}" for k, v, g in seq]
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1], per_launch="--all" in sys.argv))
