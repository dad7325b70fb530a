timeout 900 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | grep "^E  \|^FAILED\|Error\|passed\|failed" | cut -c1-300 | head -30
for i in 1 2; do
B200SR_IMG_DEPS=0 python tools/quick_bench.py 2>&1 | grep "fwd\|step" | sed 's/^/group /'
python tools/quick_bench.py 2>&1 | grep "fwd\|step" | sed 's/^/image /'
done
