run() { python tools/quick_bench.py 2>&1 | grep "step" | sed "s/^/$1 /"; }
for i in 1 2; do
B200SR_LIB=$GRAFT_REPO_ROOT/sr_gan_fd_b200/libb200sr_base.so run "base(2str)  "
run "cur(4str,99)"
B200SR_LIB=$GRAFT_REPO_ROOT/sr_gan_fd_b200/libb200sr_th8.so run "tileH8      "
done
B200SR_LIB=$GRAFT_REPO_ROOT/sr_gan_fd_b200/libb200sr_th8.so timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "gradients or golden" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -2
