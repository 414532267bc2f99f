python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 900 -c 420 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/ncu_list.log 2>&1; echo rc=$?
