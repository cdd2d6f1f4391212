#!/bin/bash
# GPU job: last sanity pass over the committed build -- full GPU suite, smoke(), the default bench line
out=gpurun_out/r3final; mkdir -p $out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee $out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $out/smoke.log
python bench.py > $out/bench_default.json 2> $out/bench_default.err; tail -c 600 $out/bench_default.json | head -c 600; echo
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; cut -c1-300 $out/bench_reference.json
