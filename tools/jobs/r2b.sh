#!/bin/bash
# GPU job: tests + small-span benches + drop-in band sweep + default bench (round 2, run b)
out=gpurun_out/r2b; mkdir -p $out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > $out/tests.log
for w in 1080p_16x16_pm1 1080p_16x16_pm2 1080p_16x16_pm4 4k_16x16_pm2; do
  python bench.py --workload $w --no-cpu-baseline --sustained-s 1 --dropin-calls 0 > $out/bench_$w.json 2> $out/bench_$w.err
done
(for r in 1 2 3 4; do python tools/quick_bench.py 1920 1080 16 $r 64; done
 for sb in 1 2 4 8 17 34 68; do echo sb=$sb; ME_B200_STREAM_SB=$sb python tools/quick_bench.py 1920 1080 16 2 64; done
 ME_B200_NO_STREAM=1 python tools/quick_bench.py 1920 1080 16 2 64) > $out/quick.log 2>&1
(for b in 1 2 3 4 6 8; do echo bands=$b; ME_B200_DROPIN_BANDS=$b python tools/dropin_latency.py; done) > $out/dropin.log 2>&1
python bench.py > $out/bench.json 2> $out/bench.err
tail -3 $out/tests.log; tail -30 $out/quick.log
