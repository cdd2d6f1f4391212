#!/bin/bash
# GPU job (late round 2): final single-GPU evidence of the round -- full GPU test suite, smoke(), bench lines of every
# workload, launch list + full ncu capture of the headline command, integer peaks
out=gpurun_out/r3z; mkdir -p $out/bench
(python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $out/tests.log; tail -3 $out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
python bench.py > $out/bench/1080p_16x16_pm32.json 2> $out/bench/1080p_16x16_pm32.err
python bench.py --impl reference > $out/bench/reference_arm.json 2> $out/bench/reference_arm.err
for w in 1080p_16x16_pm64 4k_8x8_pm32 4k_16x16_pm32 foreman_8x8_pm12 4k_8x8_pm12; do
  python bench.py --workload $w --sustained-s 1 > $out/bench/$w.json 2> $out/bench/$w.err; done
for w in 1080p_16x16_pm0 4k_16x16_pm0 1080p_16x16_pm1 1080p_16x16_pm2 1080p_16x16_pm4 4k_16x16_pm2 ssim_4k_16x16_pm7 ssim_1080p_16x16_pm32 tss_1080p_16x16_pm32 diamond_1080p_16x16_pm32 diamond_foreman_8x8_pm12; do
  python bench.py --workload $w --sustained-s 1 --no-cpu-baseline > $out/bench/$w.json 2> $out/bench/$w.err; done
python tools/int_peaks.py > $out/int_peaks.json 2>&1
CMD="python bench.py --steps 2 --warmup 3 --pairs 8 --no-cpu-baseline --sustained-s 0 --dropin-calls 0 --no-post --no-parity-check"
$CMD > $out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1
$CMD > $out/plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tiled_search -s 3 -c 1 -o $out/prof_tiled_final $CMD > $out/ncu_full.log 2>&1
python tools/quick_bench.py 3840 2160 16 1 32 > $out/plain_r1_4k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_search -s 3 -c 1 -o $out/prof_stream_pm1_4k python tools/quick_bench.py 3840 2160 16 1 32 > $out/ncu_r1.log 2>&1
python tools/quick_bench.py 3840 2160 8 12 4 > $out/plain_8x8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tiled_search -s 3 -c 1 -o $out/prof_tiled_8x8_pm12 python tools/quick_bench.py 3840 2160 8 12 4 > $out/ncu_8x8.log 2>&1
for f in $out/bench/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',r.get('frac'),'parity',d.get('parity_checked'),'dropin',(d.get('e2e_dropin') or {}).get('ms_per_call'))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
done
