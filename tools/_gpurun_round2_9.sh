# ring-depth rule for lone CTAs + per-layer plan sweep on the whole 512x512 step
mkdir -p gpurun_out
for c in conv2 conv_compact; do timeout 600 python tools/gpu_selftest.py --case $c > gpurun_out/r2_selftest_${c}_9.log 2>&1; echo "$c rc=$? pass=$(grep -c ^PASS gpurun_out/r2_selftest_${c}_9.log) fail=$(grep -c ^FAIL gpurun_out/r2_selftest_${c}_9.log)"; done
python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_bench512_v9.json 2> gpurun_out/r2_bench512_v9.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in reversed(open("gpurun_out/r2_bench512_v9.json").read().strip().splitlines()):
    if ln.startswith("{"):
        d = json.loads(ln); print("512 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["achieved"],1)); break
PY
for grp in 0,1,2 3,4,5,6 7,8,9 10,11; do
timeout 500 python tools/plan_sweep.py --size 512 --layers $grp > gpurun_out/r2_plan_sweep_512_$grp.log 2>&1; echo "sweep $grp rc=$?"; cat gpurun_out/r2_plan_sweep_512_$grp.log
done
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_512_step_warm.csv python tools/profile_step.py --size 512 --steps 1 > gpurun_out/ncu_512w.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_512_step_warm.csv | head -30
