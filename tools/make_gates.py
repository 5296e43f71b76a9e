"""Regenerate the per-fixture GATES table of tests/test_gpu_parity.py from parity reports.

    python tools/make_gates.py profiles/r2_parity_report_v1.log profiles/r2_parity_report_v3.log

gate = 1.5 x the largest deviation any given report measured (eager and graph runs), two
significant digits, never below the floor of the metric.  Several reports = the same fixtures
measured with kernels that differ only in fp32 summation order (tile families, split-K): the spread
between them is the noise a gate has to tolerate.
"""
from __future__ import annotations

import re
import sys
from collections import defaultdict

FLOORS = {"layer_style_rel_max": 5e-4, "grad_rel_l2": 2e-3, "total_rel_max": 5e-4,
          "style_rel_max": 5e-4, "final_rel_l2": 1e-4, "frames_max_lsb": 1, "frames_frac_diff": 0.01}
SKIP = {"grad_cos_min", "layer_content_abs_max", "grad_cos", "content_abs_max"}


def two_digits(v: float) -> float:
    return float(f"{v:.2g}")


def main() -> None:
    worst: dict[str, dict[str, float]] = defaultdict(dict)
    for path in sys.argv[1:]:
        for line in open(path):
            m = re.match(r"(adam_\w+) graph=(True|False): (.*)", line)
            if not m:
                continue
            for k, v in re.findall(r"(\w+)=([0-9.eE+-]+|nan)", m.group(3)):
                if k in SKIP or k not in FLOORS:
                    continue
                worst[m.group(1)][k] = max(worst[m.group(1)].get(k, 0.0), float(v))
    print("GATES: dict[str, dict[str, float]] = {")
    for name in sorted(worst):
        items = []
        for k, v in worst[name].items():
            g = max(1.5 * v, FLOORS[k])
            # an integer count of LSBs: measured + 1 (the global ceiling of 3 still applies)
            g = int(v) + 1 if k == "frames_max_lsb" else two_digits(g)
            items.append(f'"{k}": {g}')
        print(f'    "{name}": {{' + ", ".join(items) + "},")
    print("}")


if __name__ == "__main__":
    main()
