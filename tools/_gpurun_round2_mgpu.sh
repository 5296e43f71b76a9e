# usage: bash tools/_gpurun_round2_mgpu.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29611 tools/sharded_check.py --golden4k > gpurun_out/r2_sharded_check_${N}gpu.log 2>&1; echo "sharded_check rc=$?"; grep "PASS\|FAIL\|Error\|watchdog\|symmetric" gpurun_out/r2_sharded_check_${N}gpu.log | cut -c1-300 | head -20; tail -5 gpurun_out/r2_sharded_check_${N}gpu.log | cut -c1-300
for mode in peer nccl; do
  if [ $mode = nccl ]; then export STV_HALO=nccl; else unset STV_HALO; fi
  timeout 600 $TR --master-port 29612 bench.py --gpus $N --workload 4k --steps 10 --warmup 3 > gpurun_out/r2_bench_4k_${N}gpu_${mode}.json 2> gpurun_out/r2_bench_4k_${N}gpu_${mode}.err; echo "bench4k $mode rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2_bench_4k_${N}gpu_${mode}.json") if l.startswith("{")][-1]
    print("4k x$N $mode:", round(d["value"],2), "steps/s", d["config"]["halo_exchange"], "launches/step", d["gpu_launches_per_step"], "parity", d["parity_vs_golden"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2_bench_4k_${N}gpu_${mode}.err").read()[-2000:])
PY
done
unset STV_HALO
timeout 900 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_default_${N}gpu.json 2> gpurun_out/r2_bench_default_${N}gpu.err; echo "bench default rc=$?"
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2_bench_default_${N}gpu.json") if l.startswith("{")][-1]
    print("default x$N:", round(d["value"],1), "e2e", round(d["e2e"]["value"],1)); print("  sharded_4k:", {k:v for k,v in d["workloads"]["sharded_4k"].items() if k in ("value","efficiency_vs_n1","n1_value_unsharded","parity_vs_golden","halo_exchange")})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2_bench_default_${N}gpu.err").read()[-2000:])
PY
