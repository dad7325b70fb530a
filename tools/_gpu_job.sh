timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2i_bench_n2.json
python -c "
import json
d=json.loads(open('gpurun_out/r2i_bench_n2.json').read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','dp_equiv_rel_l2')}, d['e2e']['value'])"
