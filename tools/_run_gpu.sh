mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/dp_check.py > gpurun_out/r2/dp_check_n2.log 2>&1; tail -2 gpurun_out/r2/dp_check_n2.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N1', round(d['ms_per_step'],3), round(d['value'],1))"
for ov in 0 1 0 1; do
B200SR_DP_OVERLAP=$ov $TR bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2/n2_ov$ov.err | tail -1 > gpurun_out/r2/n2_ov$ov.json
python -c "import json; d=json.loads(open('gpurun_out/r2/n2_ov$ov.json').read()); print('N2 overlap=$ov', round(d['ms_per_step'],3), round(d['value'],1), d.get('dp_equiv_rel_l2'), d['gpu_launches'])"
done
B200SR_DP_OVERLAP=1 B200SR_DP_MAX_CTAS=8 $TR bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N2 overlap=1 ctas8', round(d['ms_per_step'],3), round(d['value'],1))"
B200SR_DP_OVERLAP=1 B200SR_DP_BUCKET_RRDBS=3 $TR bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N2 overlap=1 rrdbs3', round(d['ms_per_step'],3), round(d['value'],1))"
