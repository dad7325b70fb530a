timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','step_frac_of_peak','infer_out_mpix_per_s','vs_library_best')}, d['roofline']['frac'], d['roofline']['ms_forward'], d['e2e']['value'], d['e2e']['sync_readback']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['widened']['bsrgan_gan_step']['img_per_s'], d['clocks'])
PY
