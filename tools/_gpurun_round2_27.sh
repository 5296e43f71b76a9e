mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runner.py -m gpu -q -x > gpurun_out/r2_pytest_27.log 2>&1; tail -3 gpurun_out/r2_pytest_27.log
python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2_bench512_v27.json 2> gpurun_out/r2_bench512_v27.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in reversed(open("gpurun_out/r2_bench512_v27.json").read().strip().splitlines()):
    if ln.startswith("{"):
        d = json.loads(ln); print("512 (steps 20) value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"]["phases"], d["e2e"]["pass_totals_s"]); break
PY
