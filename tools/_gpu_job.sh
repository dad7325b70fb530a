timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','step_tflops','step_frac_of_peak','infer_out_mpix_per_s','vs_library_best','gpu_launches')}, d['roofline']['frac'], d['roofline']['ms_forward'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['sync_readback']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['widened']['bsrgan_gan_step']['img_per_s'], d['clocks'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches_raw.csv python tools/profile_step.py > /dev/null 2>&1; tail -1 gpurun_out/r2i_launches_raw.csv | cut -c1-100
