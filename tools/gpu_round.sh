#!/bin/bash
# One GPU-box pass: parity tests, bench line, per-layer profile, ncu launch list, ncu --set full of the top kernels.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
HCUNET_PROFILE_OUT=$OUT/layers.txt python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err; cat $OUT/bench_ref.json
if [ "$2" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_mma_kernel" -c 6 -f -o $OUT/full_$TAG \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
