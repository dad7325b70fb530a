mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu > gpurun_out/r2/pytest_gpu10.log 2>&1
tail -6 gpurun_out/r2/pytest_gpu10.log
for i in 1 2; do
for v in 0 1; do
  B200SR_K32=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('k32=$v', round(d['ms_per_step'],3), round(d['roofline']['ms_forward'],3), round(d['infer_out_mpix_per_s'],1))"
done
done > gpurun_out/r2/ab5.log 2>&1
cat gpurun_out/r2/ab5.log
