python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])"
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -1
