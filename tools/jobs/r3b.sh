#!/bin/bash
# GPU job: SSIM FORM 4 -- where does the time go?  drain statistics, perfect-pruning ceiling, fuzz error text
out=gpurun_out/r3b; mkdir -p $out
(timeout 600 python tools/fuzz_parity.py 60 11 ssim 2>&1 | tail -12) > $out/fuzz.log; cat $out/fuzz.log
w=ssim_1080p_16x16_pm32
run() { tag=$1; shift; env "$@" python bench.py --workload $w --no-cpu-baseline --no-parity-check --no-post --sustained-s 0 --steps 5 > $out/$tag.json 2> $out/$tag.err; python - <<PY
import json
for l in open("$out/$tag.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$tag", "value", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 4))
PY
}
run base X=1
run fake_thr ME_B200_SSIM_FAKE_THR=0.9999
run parts2 ME_B200_PARTS=2
run parts4 ME_B200_PARTS=4
run ns4 ME_B200_NS=4
run old ME_B200_SSIM_FORM4=0
ME_B200_SSIM_STATS=1 python bench.py --workload $w --no-cpu-baseline --no-parity-check --no-post --sustained-s 0 --steps 1 --pairs 4 > $out/stats.json 2> $out/stats.err
grep "ssim form 4" $out/stats.err | tail -3
