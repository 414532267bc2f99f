#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/cfg_bench.py cfg3 --steps 3 --out $OUT/r03c_cfg3_layers.txt > $OUT/r03c_cfg3.json 2> $OUT/r03c_cfg3.err; echo "cfg3 rc=$?"
python -c "import json;d=json.load(open('$OUT/r03c_cfg3.json'));print(d['ms_per_step'], d['train_tflops'], d['kernel_ms_per_step'])"
head -24 $OUT/r03c_cfg3_layers.txt | cut -c1-110
