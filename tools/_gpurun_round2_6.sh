mkdir -p gpurun_out
for c in elementwise conv_compact; do
timeout 900 python tools/gpu_selftest.py --case $c > gpurun_out/r2_selftest_$c.log 2>&1; echo "$c rc=$? pass=$(grep -c PASS gpurun_out/r2_selftest_$c.log)"; grep "FAIL\|Error\|watchdog" gpurun_out/r2_selftest_$c.log | head -8
done
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r2_pytest_gpu_6.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu_6.log
python bench.py --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench512_v6.json 2> gpurun_out/r2_bench512_v6.err
python bench.py --workload 1080p --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench1080_v6.json 2> gpurun_out/r2_bench1080_v6.err
python - <<'PY'
import json
def last_json(path):
    for ln in reversed(open(path).read().strip().splitlines()):
        if ln.startswith('{'): return json.loads(ln)
for f in ("r2_bench512_v6","r2_bench1080_v6"):
    try:
        d=last_json(f"gpurun_out/{f}.json")
        print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "launches/step", d["gpu_launches_per_step"], d["clocks"]["reasons"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
for sz in 512 1080p; do
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"conv_igemm2|gram_partial|conv_first|relu_fwd" --csv --log-file gpurun_out/r2_ncu_conv_gram_metrics_${sz}_v2.csv python tools/profile_step.py --size $sz --steps 1 > gpurun_out/ncu_m${sz}.log 2>&1
python tools/summarize_metrics.py gpurun_out/r2_ncu_conv_gram_metrics_${sz}_v2.csv
done
