#!/bin/bash
# GPU job (8 GPUs): name the limiter of the end-to-end arm -- copy-only H2D probe, then bench.py at N=8
# with / without the per-rank CPU affinity and with write-combined frame buffers
out=gpurun_out/r2c; mkdir -p $out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > $out/topo.txt 2>&1
lscpu > $out/lscpu.txt 2>&1
numactl -H > $out/numa.txt 2>&1
$TR tools/h2d_probe.py > $out/h2d_probe_n$N.json 2> $out/h2d_probe_n$N.err
$TR bench.py --gpus $N --no-cpu-baseline > $out/bench_n$N.json 2> $out/bench_n$N.err
BENCH_AFFINITY=0 $TR bench.py --gpus $N --no-cpu-baseline --no-band-split --sustained-s 0 --dropin-calls 0 --no-post > $out/bench_n${N}_noaff.json 2> $out/bench_n${N}_noaff.err
BENCH_WC=1 $TR bench.py --gpus $N --no-cpu-baseline --no-band-split --sustained-s 0 --dropin-calls 0 --no-post > $out/bench_n${N}_wc.json 2> $out/bench_n${N}_wc.err
python bench.py --no-cpu-baseline --no-post --sustained-s 0 --dropin-calls 0 > $out/bench_n1.json 2> $out/bench_n1.err
for f in $out/bench_n*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['e2e']['step_synchronous']),'h2d/gpu',round(d['e2e']['h2d_gbs_per_gpu'],1),'seq',round(d['e2e_sequence']['value']))
"; done
