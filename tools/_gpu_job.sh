timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gan_step_native_gpu.py -m gpu -q -s 2>&1 | grep "all-native" | cut -c1-300
