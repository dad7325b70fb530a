mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu > gpurun_out/r2/pytest_gpu9.log 2>&1
tail -3 gpurun_out/r2/pytest_gpu9.log
for i in 1 2; do
for v in base cur; do
  if [ $v = base ]; then export B200SR_LIB=$GRAFT_REPO_ROOT/sr_gan_fd_b200/libb200sr_base.so; else unset B200SR_LIB; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['roofline']['ms_forward'],3))"
done
done > gpurun_out/r2/ab4.log 2>&1
cat gpurun_out/r2/ab4.log
