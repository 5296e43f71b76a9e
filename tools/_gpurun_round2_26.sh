mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "non_finite or warnings or runner or hooks" > gpurun_out/r2_pytest_26.log 2>&1; tail -3 gpurun_out/r2_pytest_26.log
python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_bench512_v26.json 2> gpurun_out/r2_bench512_v26.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in reversed(open("gpurun_out/r2_bench512_v26.json").read().strip().splitlines()):
    if ln.startswith("{"):
        d = json.loads(ln); print("512 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"]["pass_totals_s"]); break
PY
