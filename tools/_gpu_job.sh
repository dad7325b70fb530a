run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 8 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"; }
run "coalesced          " 29520
B200SR_DP_OVERLAP=1 B200SR_DP_BUCKET_RRDBS=6 run "overlap rrdb6      " 29521
B200SR_DP_OVERLAP=1 B200SR_DP_BUCKET_RRDBS=12 run "overlap rrdb12     " 29522
B200SR_DP_OVERLAP=1 B200SR_DP_BUCKET_RRDBS=12 B200SR_DP_MAX_CTAS=8 run "overlap rrdb12 cta8" 29523
