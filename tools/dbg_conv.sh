python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/kernel_bench.py conv 2>&1 | grep "^conv "
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['gpu_launches'])"
