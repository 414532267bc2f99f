python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bn_bwd_stats" --launch-skip 90 -c 45 --csv --log-file gpurun_out/l_bn.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/l_bn.csv 5
