#!/bin/bash
# GPU job: launch list of the 64-pair headline step (shares of the pre-pass and the search)
out=gpurun_out/r3f; mkdir -p $out
A="--workload 1080p_16x16_pm32 --no-cpu-baseline --no-post --sustained-s 0 --dropin-calls 0 --no-parity-check --no-band-split --steps 2 --warmup 3"
python bench.py $A > $out/plain.json 2> $out/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches.csv python bench.py $A > $out/ncu.log 2>&1
python - <<'PY'
import csv
for l in csv.reader(open("gpurun_out/r3f/launches.csv")):
    if len(l) > 14 and l[0].isdigit():
        n = l[4]
        short = "tiled" if "tiled_search" in n else "box" if "box_energy" in n else n[:30]
        print(l[0], short, l[8], l[14])
PY
