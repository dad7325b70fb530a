timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2g_c4_n1.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_c4_n1.json').read())
print({k:d.get(k) for k in ('metric','value','unit','ms_per_step','n_gpus')}); print(json.dumps(d.get('config'))[:600]); print({k:v for k,v in d.items() if k not in ('config','clocks','metric')})
PY
