#!/bin/bash
# One 8-GPU box: BASELINE configs 2-5 at their stated GPU counts.  Logs under gpurun_out/r2_mg_*.json
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out
$TR --nproc-per-node 8 --master-port 29501 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_mg_s3d_n8.json 2> $O/r2_mg_s3d_n8.err
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29510+n)) bench.py --gpus $n --encoder x3dl --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_mg_x3dl_n$n.json 2> $O/r2_mg_x3dl_n$n.err
done
$TR --nproc-per-node 8 --master-port 29530 bench.py --gpus 8 --encoder slowfast4x16 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_mg_sf_n8.json 2> $O/r2_mg_sf_n8.err
$TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --train --batch 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_mg_train_n8.json 2> $O/r2_mg_train_n8.err
for f in $O/r2_mg_*.json; do echo "== $f"; tail -c 300 $f; echo; done
for f in $O/r2_mg_*.err; do echo "== $f"; tail -n 3 $f; done
