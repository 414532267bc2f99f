python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for n in 1 2; do HCUNET_SIDE_STREAMS=$n python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('side streams $n: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['gpu_launches'])"; done
