#!/bin/bash
# GPU job (2 GPUs): peer / band-sharding / ingest tests with the final kernels (8x8 peer instantiations use the stage queue)
out=gpurun_out/r3u; mkdir -p $out
(timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -6) | tee $out/tests_peer.log
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ingest" 2>&1 | tail -4) | tee $out/tests_ingest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-cpu-baseline --sustained-s 0 --dropin-calls 0 --no-post > $out/bench_n2.json 2> $out/bench_n2.err
python - <<'PY'
import json
for l in open("gpurun_out/r3u/bench_n2.json"):
    if l.startswith("{"):
        d = json.loads(l); print("n2 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_checked"), "band_split", d.get("band_split"))
PY
tail -3 $out/bench_n2.err
