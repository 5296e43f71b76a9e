mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_pytest_30.log 2>&1; tail -2 gpurun_out/r2_pytest_30.log
python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2_bench512_v30.json 2> gpurun_out/r2_bench512_v30.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in reversed(open("gpurun_out/r2_bench512_v30.json").read().strip().splitlines()):
    if ln.startswith("{"):
        d = json.loads(ln); print("512 (steps 20) value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1)); break
PY
