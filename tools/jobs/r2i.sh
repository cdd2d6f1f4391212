#!/bin/bash
# GPU job: arriving-frame drop-in -- parity + latency
out=gpurun_out/r2i; mkdir -p $out
(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "drop_in" 2>&1 | tail -30) > $out/tests.log
tail -4 $out/tests.log
(echo arrive default; python tools/dropin_latency.py
 for b in 4 8 16; do echo arrive bands=$b; ME_B200_DROPIN_BANDS=$b python tools/dropin_latency.py; done
 echo banded; ME_B200_DROPIN_ARRIVE=0 python tools/dropin_latency.py
 ME_B200_TRACE=1 python tools/dropin_latency.py 2>&1 | grep "bands 8" | tail -2) > $out/dropin.log 2>&1
cat $out/dropin.log
python bench.py --no-cpu-baseline --sustained-s 0 > $out/bench.json 2> $out/bench.err; tail -3 $out/bench.err
python -c "
import json
d=json.load(open('$out/bench.json')); print('value',d['value'],'e2e',d['e2e']['value'],'dropin',d['e2e_dropin'])"
