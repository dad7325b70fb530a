run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline --no-widened 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(d['roofline']['ms_forward'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"; }
for i in 1 2; do
  export B200SR_LIB=$GRAFT_REPO_ROOT/sr_gan_fd_b200/libb200sr_base.so; run base
  unset B200SR_LIB; run cur
done
timeout 600 python -m pytest tests/test_vgg.py tests/test_model_gpu.py -q -m gpu -k "vgg or gpu_ or full_config2 or golden" 2>&1 | tail -2
