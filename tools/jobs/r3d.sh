#!/bin/bash
# GPU job (8 GPUs): ingest-helper routing with the conservative rule (detour <= 80 % of the donor's spare), and 2 pairs forced
out=gpurun_out/r3d; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
A="--gpus 8 --no-cpu-baseline --no-band-split --sustained-s 0 --dropin-calls 0 --no-post --no-parity-check"
$TR --nproc-per-node 8 --master-port 29531 bench.py $A > $out/bench_n8_rule.json 2> $out/bench_n8_rule.err
BENCH_INGEST_HP=2 $TR --nproc-per-node 8 --master-port 29532 bench.py $A > $out/bench_n8_hp2.json 2> $out/bench_n8_hp2.err
BENCH_INGEST_HELPER=0 $TR --nproc-per-node 8 --master-port 29533 bench.py $A > $out/bench_n8_nohelper.json 2> $out/bench_n8_nohelper.err
for f in $out/bench_n8_*.json; do echo $f; python -c "
import json
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['e2e']['step_synchronous']),'h2d/gpu',round(d['e2e']['h2d_gbs_per_gpu'],1),'ingest',(d['e2e'].get('ingest_routing') or {}).get('helpers'))
"; tail -2 ${f%.json}.err; done
