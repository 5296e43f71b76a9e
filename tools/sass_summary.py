"""Count the Blackwell tensor-core / TMEM / TMA instructions per kernel of libstv_b200.so
(`cuobjdump -sass`), as evidence that the hot kernels are tcgen05 / TMEM / TMA code.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
from __future__ import annotations

import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "style_transfer_visualizer_b200" / "lib" / "libstv_b200.so"
MNEMONICS = OrderedDict([
    ("UTCHMMA", "tcgen05.mma (tf32 / f16 kinds)"),
    ("UTCHMMA.2CTA", "  of which cta_group::2"),
    ("LDTM", "tcgen05.ld (TMEM -> registers)"),
    ("UTMALDG", "TMA tensor load (cp.async.bulk.tensor)"),
    ("UTMASTG", "TMA tensor store"),
    ("UTCBAR", "tcgen05.commit (mbarrier arrive)"),
    ("SYNCS", "mbarrier ops"),
    ("UTCATOMSWS", "tcgen05.alloc / dealloc"),
    ("STG.E.128", "16-byte global stores"),
    ("LDL", "local-memory loads (spills / printf args)"),
])


def main() -> None:
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True,
                         check=True).stdout
    demangle = subprocess.run(["cu++filt"], input="\n".join(
        re.findall(r"Function : (\S+)", out)), capture_output=True, text=True, check=False).stdout
    names = dict(zip(re.findall(r"Function : (\S+)", out), demangle.splitlines()))
    counts: dict[str, dict[str, int]] = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names.get(m.group(1), m.group(1))
            cur = cur.replace("(int)", "").replace("(bool)", "")
            cur = re.sub(r"\(.*", "", cur).replace("void stv::", "").replace("stv::", "")
            counts[cur] = {k: 0 for k in MNEMONICS}
            continue
        if cur is None:
            continue
        mm = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not mm:
            continue
        op = mm.group(1)
        for key in MNEMONICS:
            if key == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    counts[cur][key] += 1
            elif op.startswith(key):
                counts[cur][key] += 1
    print(f"# {LIB.relative_to(ROOT)}: instruction counts per kernel (cuobjdump -sass, sm_100a)")
    for key, what in MNEMONICS.items():
        print(f"#   {key:<14s} {what}")
    keys = list(MNEMONICS)
    print(f"{'kernel':<58s} " + " ".join(f"{k[:9]:>9s}" for k in keys))
    total = {k: 0 for k in keys}
    for name, c in counts.items():
        if not any(c.values()):
            continue
        print(f"{name[:58]:<58s} " + " ".join(f"{c[k]:9d}" for k in keys))
        for k in keys:
            total[k] += c[k]
    print(f"{'TOTAL':<58s} " + " ".join(f"{total[k]:9d}" for k in keys))
    sys.stdout.flush()


if __name__ == "__main__":
    main()
