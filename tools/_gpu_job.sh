timeout 900 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py tests/test_disc.py tests/test_vgg.py -m gpu -q -x 2>&1 | tail -2
