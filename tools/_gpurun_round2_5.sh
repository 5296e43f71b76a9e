mkdir -p gpurun_out
for c in halo conv_compact conv_variants conv2; do
timeout 900 python tools/gpu_selftest.py --case $c > gpurun_out/r2_selftest_$c.log 2>&1; echo "$c rc=$? pass=$(grep -c PASS gpurun_out/r2_selftest_$c.log)"; grep "FAIL\|Error\|watchdog" gpurun_out/r2_selftest_$c.log | head -8
done
timeout 900 python tools/epilogue_ab.py > gpurun_out/r2_epilogue_ab.log 2>&1; cat gpurun_out/r2_epilogue_ab.log | tail -70
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r2_pytest_gpu_5.log 2>&1; tail -8 gpurun_out/r2_pytest_gpu_5.log
python bench.py --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench512_v5.json 2> gpurun_out/r2_bench512_v5.err
python bench.py --workload 1080p --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench1080_v5.json 2> gpurun_out/r2_bench1080_v5.err
python - <<'PY'
import json
for f in ("r2_bench512_v5","r2_bench1080_v5"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "launches/step", d["gpu_launches_per_step"], d["clocks"]["reasons"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
