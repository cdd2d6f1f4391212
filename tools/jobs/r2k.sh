#!/bin/bash
# GPU job: small-span workloads with the new stripe model and batch sizes
out=gpurun_out/r2k; mkdir -p $out/bench
(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_span" 2>&1 | tail -3) > $out/tests.log; cat $out/tests.log
(for r in 1 2 3 4; do for np in 1 8 64 256; do python tools/quick_bench.py 1920 1080 16 $r $np; done; done
 python tools/quick_bench.py 3840 2160 16 1 64; python tools/quick_bench.py 3840 2160 16 2 64) > $out/quick.log 2>&1; cat $out/quick.log
for w in 1080p_16x16_pm1 1080p_16x16_pm2 1080p_16x16_pm4 4k_16x16_pm2; do
  python bench.py --workload $w --sustained-s 1 --no-cpu-baseline > $out/bench/$w.json 2> $out/bench/$w.err; tail -2 $out/bench/$w.err
  python -c "
import json
d=json.load(open('$out/bench/$w.json')); r=d['roofline']; print('$w','value',round(d['value']),'frac',r['bound'],round(r['frac'],3), r.get('int_alu'), 'e2e', round(d['e2e']['value']),'parity',d['parity_checked'])"; done
