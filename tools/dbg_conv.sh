python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/parity_probe.py 2>&1 | grep -v Warn | tail -10
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['kernel_time_shares'])"
tail -3 gpurun_out/bench_quick.err
