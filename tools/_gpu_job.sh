timeout 300 python -m pytest tests/test_disc.py -m gpu -q -s 2>&1 | grep "^disc (\|passed\|failed\|FAILED\|^E " | cut -c1-400
timeout 300 python tools/bench_disc.py 2>&1 | tail -1 | tee gpurun_out/bench_disc6_fp16.json | cut -c1-330
timeout 300 python tools/gan_step.py --steps 10 --disc b200 --content b200 --optim fused --ema 2>/dev/null | tail -1 | cut -c150-420
timeout 600 python bench.py --no-cpu-baseline --no-library-baseline > gpurun_out/bench_widened.json 2>gpurun_out/bench_widened.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_widened.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('widened'))[:1800])
"
