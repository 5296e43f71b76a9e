mkdir -p gpurun_out
for c in conv_compact elementwise; do
timeout 900 python tools/gpu_selftest.py --case $c > gpurun_out/r2_selftest_$c.log 2>&1; echo "$c rc=$? pass=$(grep -c PASS gpurun_out/r2_selftest_$c.log)"; grep "FAIL\|Error\|watchdog\|tensor-core" gpurun_out/r2_selftest_$c.log | head -20
done
timeout 600 python tools/gpu_parity_report.py > gpurun_out/r2_parity_report_2.log 2>&1; grep "graph=" gpurun_out/r2_parity_report_2.log | cut -c1-330
python bench.py --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench512_v8.json 2> gpurun_out/r2_bench512_v8.err
python bench.py --workload 1080p --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench1080_v8.json 2> gpurun_out/r2_bench1080_v8.err
python - <<'PY'
import json
def last_json(path):
    for ln in reversed(open(path).read().strip().splitlines()):
        if ln.startswith('{'): return json.loads(ln)
for f in ("r2_bench512_v8","r2_bench1080_v8"):
    try:
        d=last_json(f"gpurun_out/{f}.json")
        print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "launches/step", d["gpu_launches_per_step"], d["clocks"]["reasons"], d["clocks"]["sm_mhz"])
        for r in d["roofline_hbm"]: print("   hbm", r["kernel"], round(r["achieved"]), round(r["frac"],2))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
