run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(d['roofline']['ms_forward'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"; }
for i in 1 2; do
  unset B200SR_LIB; run cur
  for v in f1 f2 p1; do export B200SR_LIB=$GRAFT_REPO_ROOT/sr_gan_fd_b200/libb200sr_$v.so; run $v; done
  unset B200SR_LIB; B200SR_TAPS1=1 run taps1
done
B200SR_TAPS1=1 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "full_config2 or golden" 2>&1 | tail -2
B200SR_LIB=$GRAFT_REPO_ROOT/sr_gan_fd_b200/libb200sr_f2.so python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "full_config2 or golden" 2>&1 | tail -2
