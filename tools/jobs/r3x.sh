#!/bin/bash
# GPU job: refresh the 8x8 bench lines and the 8x8 ncu capture after the epilogue change
out=gpurun_out/r3x; mkdir -p $out/bench
for w in foreman_8x8_pm12 4k_8x8_pm12 4k_8x8_pm32; do
  python bench.py --workload $w --sustained-s 1 > $out/bench/$w.json 2> $out/bench/$w.err; done
python tools/quick_bench.py 3840 2160 8 12 4 > $out/plain_8x8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tiled_search -s 3 -c 1 -o $out/prof_tiled_8x8_pm12 python tools/quick_bench.py 3840 2160 8 12 4 > $out/ncu_8x8.log 2>&1
for f in $out/bench/*.json; do python - "$f" <<'PY'
import json,sys
d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][0]
print(sys.argv[1].split('/')[-1], 'value', round(d['value'],1), 'frac', round(d['roofline']['frac'],4), 'parity', d.get('parity_checked'), 'dropin', (d.get('e2e_dropin') or {}).get('ms_per_call'))
PY
done
