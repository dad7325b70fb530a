mkdir -p gpurun_out/r2
python -m pytest tests/test_model_gpu.py -x -q -m gpu > gpurun_out/r2/pytest_gpu3.log 2>&1
tail -3 gpurun_out/r2/pytest_gpu3.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2/bench3.json 2> gpurun_out/r2/bench3.err
cat gpurun_out/r2/bench3.json | head -c 300
LO=40 HI=72 NCP=1 python tools/probe_timeline.py > gpurun_out/r2/timeline3.log 2>&1
