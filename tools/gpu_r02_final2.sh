#!/bin/bash
# Round-2 evidence pass after the row-stacked weight gradient (wgrad_rows.cu): full GPU suite, smoke, bench (ours / reference /
# fp32), timeline, then ONE ncu session (launch list of a bench step, --set full of wgrad_rows at the bench shapes, DRAM traffic
# of the conv_tc launches of one step).
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; grep "^smoke" $OUT/${TAG}_smoke.log
HCUNET_PROFILE_OUT=$OUT/${TAG}_layers.txt timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; echo "bench ref rc=$?"
python - <<PY
import json
for f in ("${TAG}_bench", "${TAG}_bench_ref"):
    try:
        d = json.load(open("$OUT/" + f + ".json"))
        print(f, d.get("ms_per_step"), (d.get("e2e") or {}).get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("value"))
    except Exception as e:
        print(f, "no json", e)
PY
timeout 120 python tools/step_timeline.py > $OUT/${TAG}_timeline.txt 2>/dev/null; grep "^#" $OUT/${TAG}_timeline.txt | head -3
timeout 120 python tools/step_breakdown.py > $OUT/${TAG}_step_breakdown.txt 2>/dev/null; head -5 $OUT/${TAG}_step_breakdown.txt
L="d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 d2.conv2 d3.conv1 u3.conv1"
timeout 120 python tools/kernel_bench.py wgradrows $L > $OUT/${TAG}_wgrad_rows_layers.txt 2>&1
timeout 120 python tools/kernel_bench.py wgrad_auto d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 u3.conv1 >> $OUT/${TAG}_wgrad_rows_layers.txt 2>&1
timeout 120 python tools/kernel_bench.py wgrad5 d2.conv2 d3.conv1 >> $OUT/${TAG}_wgrad_rows_layers.txt 2>&1
cat $OUT/${TAG}_wgrad_rows_layers.txt
if [ "$2" == "noncu" ]; then exit 0; fi
timeout 120 python tools/kernel_bench.py wgradrows $L --once > $OUT/${TAG}_plain1.log 2>&1 && \
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile --no-extra > $OUT/${TAG}_plain3.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:"wgrad_rows_kernel" -c 8 -f -o $OUT/${TAG}_full_wgrad_rows \
  python tools/kernel_bench.py wgradrows $L --once > $OUT/${TAG}_ncu1.log 2>&1; echo "ncu full wgrad_rows rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile --no-extra > $OUT/${TAG}_ncu3.log 2>&1; echo "ncu list rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_tc_kernel|conv_ks_kernel" --launch-skip 102 -c 52 --csv \
  --log-file $OUT/${TAG}_conv_tc_dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-extra > $OUT/${TAG}_ncu4.log 2>&1; echo "ncu dram rc=$?"
}
ls -la $OUT/${TAG}_full_wgrad_rows.ncu-rep 2>/dev/null
