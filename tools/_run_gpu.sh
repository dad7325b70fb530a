mkdir -p gpurun_out/r2
export SRGANFD_REFERENCE=$GRAFT_REPO_ROOT/baseline/_ref
python -m pytest tests/test_reference_scripts.py -x -q -s -m gpu > gpurun_out/r2/ref_scripts_gpu.log 2>&1
python -m pytest tests -x -q -m gpu -s > gpurun_out/r2/pytest_gpu1.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench1.json 2> gpurun_out/r2/bench1.err
for ex in 0 1; do B200SR_EARLYX=$ex python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2/bench1_earlyx$ex.json 2>&1; done
tail -3 gpurun_out/r2/ref_scripts_gpu.log; tail -3 gpurun_out/r2/pytest_gpu1.log
