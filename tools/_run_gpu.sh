mkdir -p gpurun_out/r2
B200SR_VERBOSE=1 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/r2/pytest_gpu5.log 2>&1
tail -5 gpurun_out/r2/pytest_gpu5.log
for c in 1 2 4; do
B200SR_VERBOSE=1 B200SR_CLUSTER=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2/bench5_c$c.json 2> gpurun_out/r2/bench5_c$c.err
head -c 250 gpurun_out/r2/bench5_c$c.json; echo; tail -2 gpurun_out/r2/bench5_c$c.err
done
B200SR_CLUSTER=4 timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu > gpurun_out/r2/pytest_gpu5_c4.log 2>&1
tail -3 gpurun_out/r2/pytest_gpu5_c4.log
