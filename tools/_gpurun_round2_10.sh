# split-K second issuer for one-half tiles: correctness + A/B on the whole step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2_pytest_kernels_10.log 2>&1; tail -15 gpurun_out/r2_pytest_kernels_10.log
for sz in 512 1080p; do timeout 300 python tools/plan_sweep.py --size $sz --ab-split --reps $([ $sz = 512 ] && echo 200 || echo 40) 2>&1 | tee gpurun_out/r2_split_ab_$sz.log; done
python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_bench512_v10.json 2> gpurun_out/r2_bench512_v10.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in reversed(open("gpurun_out/r2_bench512_v10.json").read().strip().splitlines()):
    if ln.startswith("{"):
        d = json.loads(ln); print("512 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["achieved"],1)); break
PY
