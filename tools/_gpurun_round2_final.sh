# final single-GPU validation + evidence of the round
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_default_final.json 2> gpurun_out/r2_bench_default_final.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_driver_like.json 2> gpurun_out/r2_bench_driver_like.err; echo "bench(20,5) rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
def last_json(path):
    for ln in reversed(open(path).read().strip().splitlines()):
        if ln.startswith('{'): return json.loads(ln)
for f in ("r2_bench_default_final","r2_bench_driver_like"):
    try:
        d=last_json(f"gpurun_out/{f}.json")
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2000:]); continue
    print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "blocks", len(d["blocks_ms"]), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["clocks"]["samples"], "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2))
    r=d["roofline"]; print("   roofline", round(r["achieved"],1), round(r["frac"],3), r["launches_per_step"], round(r["share_of_step_time"],3))
    for rr in d.get("roofline_hbm") or []: print("   hbm", rr["kernel"], round(rr["achieved"]), round(rr["frac"],2), rr["fits_l2"])
    for k,v in (d.get("workloads") or {}).items():
        if "value" in v: print("   wl",k, round(v["value"],1), "e2e", round(v["e2e"]["value"],1) if v.get("e2e") else None, {kk:round(v["roofline"][kk],3) for kk in ("achieved","frac")} if v.get("roofline") else "", v.get("frames_read_back"))
        else: print("   wl",k,v.get("steps_per_s"))
        for rr in v.get("roofline_hbm") or []: print("      hbm", rr["kernel"], round(rr["achieved"]), round(rr["frac"],2))
print("reference arm:", open("gpurun_out/r2_bench_reference_arm.json").read()[:300])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench_512.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_bench_512.csv | head -24
for sz in 512 1080p; do
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_${sz}_step_v3.csv python tools/profile_step.py --size $sz --steps 1 > gpurun_out/ncu_${sz}.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_${sz}_step_v3.csv | head -26
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"conv_igemm2|gram_partial|conv_first|relu_fwd_bits" --csv --log-file gpurun_out/r2_ncu_conv_gram_metrics_${sz}_v3.csv python tools/profile_step.py --size $sz --steps 1 > gpurun_out/ncu_m${sz}.log 2>&1
python tools/summarize_metrics.py gpurun_out/r2_ncu_conv_gram_metrics_${sz}_v3.csv
done
