for i in 1 2; do
B200SR_LIB=$GRAFT_REPO_ROOT/sr_gan_fd_b200/libb200sr_base.so python tools/quick_bench.py 2>&1 | grep "step\|rror" | sed 's/^/base    /'
python tools/quick_bench.py 2>&1 | grep "step\|rror\|timeout" | sed 's/^/split   /'
done
B200SR_BWD_SPLIT=0 python tools/quick_bench.py 2>&1 | grep "step\|rror" | sed 's/^/split=0 /'
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -2
