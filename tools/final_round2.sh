#!/bin/bash
# One B200: the round's final single-GPU evidence.  Everything lands under gpurun_out/r2_final_*.
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --breakdown $O/r2_final_breakdown.json > $O/r2_final_bench.json 2> $O/r2_final_bench.err; tail -c 400 $O/r2_final_bench.json; echo
timeout 300 python bench.py --impl reference > $O/r2_final_ref.json 2> $O/r2_final_ref.err; tail -c 300 $O/r2_final_ref.json; echo
for e in x3dl slowfast4x16; do
  timeout 400 python bench.py --encoder $e --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-e2e > $O/r2_final_$e.json 2> $O/r2_final_$e.err
  python -c "import json;d=json.load(open('$O/r2_final_$e.json'));print('$e',d['value'],d['ms_per_step'],d['clocks'],d.get('parity_at_bench_config',{}).get('map_maxabs_minmax'))"
done
timeout 400 python bench.py --train --batch 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_final_train.json 2> $O/r2_final_train.err
python -c "import json;d=json.load(open('$O/r2_final_train.json'));print('train',d['value'],d['ms_per_step'],d['clocks'])"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2_final_bench_launches_raw.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-e2e --no-parity > $O/r2_final_bench_ncu.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2_final_launches_raw.csv python tools/prof_infer.py 2 > $O/r2_final_infer_ncu.log 2>&1
ls -la $O/r2_final_*
