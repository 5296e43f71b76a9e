set -x
mkdir -p gpurun_out
python tools/gpu_selftest.py --case conv_compact > gpurun_out/r2_selftest_compact.log 2>&1; echo "rc=$?" >> gpurun_out/r2_selftest_compact.log
tail -5 gpurun_out/r2_selftest_compact.log
grep -c PASS gpurun_out/r2_selftest_compact.log; grep FAIL gpurun_out/r2_selftest_compact.log | head -20
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_1.log 2>&1; tail -15 gpurun_out/r2_pytest_gpu_1.log
timeout 600 python tools/gpu_parity_report.py > gpurun_out/r2_parity_report_1.log 2>&1; grep "graph=" gpurun_out/r2_parity_report_1.log | cut -c1-330
for c in 1 0; do
python bench.py --steps 100 --warmup 20 --no-cpu-baseline --compact-backward $c > gpurun_out/r2_bench512_c$c.json 2> gpurun_out/r2_bench512_c$c.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench512_c$c.json'));print('512 compact=$c',d['value'],d['e2e']['value'],d['roofline']['achieved'],d['gpu_launches_per_step'])"
python bench.py --workload 1080p --steps 50 --warmup 10 --no-cpu-baseline --compact-backward $c > gpurun_out/r2_bench1080_c$c.json 2> gpurun_out/r2_bench1080_c$c.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench1080_c$c.json'));print('1080 compact=$c',d['value'],d['e2e']['value'],d['roofline']['achieved'],d['gpu_launches_per_step'])"
done
