for i in 1 2; do
B200SR_LIB=$GRAFT_REPO_ROOT/sr_gan_fd_b200/libb200sr_base.so python tools/quick_bench.py 2>&1 | grep "step" | sed 's/^/base   /'
python tools/quick_bench.py 2>&1 | grep "step" | sed 's/^/biasvec /'
done
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py tests/test_disc.py -m gpu -q -x 2>&1 | tail -2
