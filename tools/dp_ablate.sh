#!/bin/bash
# Attribution of the data-parallel overhead: the N-rank bench with one class of exchanges switched off at a time (results are
# wrong by construction - timing only).  usage: tools/dp_ablate.sh N  -> gpurun_out/dp_ablate_nN.txt
N=${1:-2}
OUT=gpurun_out/dp_ablate_n$N.txt
: > $OUT
PORT=29600
for AB in none small dense a2a_ids a2a_rows small,dense,a2a_ids,a2a_rows; do
  PORT=$((PORT+1))
  A=$AB; [ "$AB" = none ] && A=""
  CDCMDR_DP_ABLATE=$A timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$AB', d['ms_per_step'], d['value'])" >> $OUT
done
cat $OUT
