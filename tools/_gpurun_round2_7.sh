mkdir -p gpurun_out
timeout 600 python tools/epilogue_ab_512.py > gpurun_out/r2_epilogue_ab_512.log 2>&1; cat gpurun_out/r2_epilogue_ab_512.log | cut -c1-130
for c in conv2 conv_variants conv_compact halo elementwise; do
timeout 900 python tools/gpu_selftest.py --case $c > gpurun_out/r2_selftest_$c.log 2>&1; echo "$c rc=$? pass=$(grep -c PASS gpurun_out/r2_selftest_$c.log)"; grep "FAIL\|Error\|watchdog" gpurun_out/r2_selftest_$c.log | head -8
done
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r2_pytest_gpu_7.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_7.log
python bench.py --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench512_v7.json 2> gpurun_out/r2_bench512_v7.err
python bench.py --workload 1080p --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench1080_v7.json 2> gpurun_out/r2_bench1080_v7.err
python - <<'PY'
import json
def last_json(path):
    for ln in reversed(open(path).read().strip().splitlines()):
        if ln.startswith('{'): return json.loads(ln)
for f in ("r2_bench512_v7","r2_bench1080_v7"):
    try:
        d=last_json(f"gpurun_out/{f}.json")
        print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "launches/step", d["gpu_launches_per_step"], d["clocks"]["reasons"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
