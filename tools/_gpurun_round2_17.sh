mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2_pytest_kernels_17.log 2>&1; tail -12 gpurun_out/r2_pytest_kernels_17.log
for sz in 512 1080p; do timeout 300 python tools/plan_sweep.py --size $sz --ab-resident --reps $([ $sz = 512 ] && echo 200 || echo 40) 2>&1 | tee gpurun_out/r2_resident_ab_$sz.log; done
python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_bench512_v17.json 2> gpurun_out/r2_bench512_v17.err; echo "bench rc=$?"
python bench.py --no-extras --no-cpu-baseline --workload 1080p > gpurun_out/r2_bench1080_v17.json 2> gpurun_out/r2_bench1080_v17.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r2_bench512_v17", "r2_bench1080_v17"):
    for ln in reversed(open(f"gpurun_out/{f}.json").read().strip().splitlines()):
        if ln.startswith("{"):
            d = json.loads(ln); print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3))
            for rr in d.get("roofline_hbm") or []: print("   hbm", rr["kernel"], round(rr["achieved"]), round(rr["frac"],2))
            break
PY
