#!/bin/bash
# GPU job: full GPU suite + headline bench after the energy pre-pass change
out=gpurun_out/r3e; mkdir -p $out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $out/tests.log; cat $out/tests.log
for w in 1080p_16x16_pm32 4k_16x16_pm32 4k_8x8_pm12; do
  python bench.py --workload $w --no-cpu-baseline --no-post --sustained-s 0 --dropin-calls 0 > $out/$w.json 2> $out/$w.err
  python - <<PY
import json
for l in open("$out/$w.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$w", "value", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 4), "parity", d.get("parity_checked"))
PY
done
