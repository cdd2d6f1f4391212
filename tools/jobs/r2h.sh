#!/bin/bash
# GPU job (8 GPUs): ingest helper (host -> peer GPU -> NVLink) test + bench.py at N = 8 with and without it, N = 4, 2
out=gpurun_out/r2h; mkdir -p $out
(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ingest" 2>&1 | tail -5) > $out/tests.log; cat $out/tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline > $out/bench_n8.json 2> $out/bench_n8.err
BENCH_INGEST_HELPER=0 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --no-cpu-baseline --no-band-split --sustained-s 0 --dropin-calls 0 --no-post > $out/bench_n8_nohelper.json 2> $out/bench_n8_nohelper.err
$TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --no-cpu-baseline --no-band-split --sustained-s 0 --dropin-calls 0 --no-post > $out/bench_n4.json 2> $out/bench_n4.err
$TR --nproc-per-node 2 --master-port 29524 bench.py --gpus 2 --no-cpu-baseline --sustained-s 0 --dropin-calls 0 --no-post > $out/bench_n2.json 2> $out/bench_n2.err
python bench.py --no-cpu-baseline --no-post --sustained-s 0 --dropin-calls 0 > $out/bench_n1.json 2> $out/bench_n1.err
for f in $out/bench_n*.json; do echo $f; python -c "
import json
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['e2e']['step_synchronous']),'h2d/gpu',round(d['e2e']['h2d_gbs_per_gpu'],1),'seq',round(d['e2e_sequence']['value']),'ingest',d['e2e'].get('ingest_routing'))
"; tail -2 ${f%.json}.err; done
