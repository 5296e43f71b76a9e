mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_3.log 2>&1; tail -25 gpurun_out/r2_pytest_gpu_3.log
timeout 300 python tools/gpu_parity_report.py lbfgs_noisy_64 lbfgs_random_64 > gpurun_out/r2_parity_lbfgs.log 2>&1; grep -A3 "graph=" gpurun_out/r2_parity_lbfgs.log | cut -c1-300
python bench.py > gpurun_out/r2_bench_default_v1.json 2> gpurun_out/r2_bench_default_v1.err; tail -3 gpurun_out/r2_bench_default_v1.err
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_s20_$i.json 2> gpurun_out/r2_bench_s20_$i.err; done
python - <<'PY'
import json
for f in ("r2_bench_default_v1","r2_bench_s20_1","r2_bench_s20_2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "blocks", len(d["blocks_ms"]), [round(b,2) for b in d["blocks_ms"][:6]], "clocks", d["clocks"])
    print("   roofline", {k:d["roofline"][k] for k in ("achieved","peak","frac","launches_per_step","share_of_step_time")})
    for r in d.get("roofline_hbm") or []: print("   hbm", r["kernel"], round(r["achieved"]), round(r["frac"],2), r["fits_l2"])
    for k,v in (d.get("workloads") or {}).items():
        if "value" in v: print("   wl",k, round(v["value"],1), "e2e", round(v["e2e"]["value"],1) if v.get("e2e") else None, {kk:round(v["roofline"][kk],3) for kk in ("achieved","frac")} if v.get("roofline") else "")
        else: print("   wl",k,v.get("steps_per_s"))
        for r in v.get("roofline_hbm") or []: print("      hbm", r["kernel"], round(r["achieved"]), round(r["frac"],2))
PY
