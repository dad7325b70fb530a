mkdir -p gpurun_out/r2
export SRGANFD_REFERENCE=$GRAFT_REPO_ROOT/baseline/_ref
python tools/bench_vgg.py > gpurun_out/r2/bench_vgg.json 2> gpurun_out/r2/bench_vgg.err; cat gpurun_out/r2/bench_vgg.json; tail -3 gpurun_out/r2/bench_vgg.err
timeout 600 python tools/gan_step.py --generator b200 --content b200 --steps 10 2> gpurun_out/r2/gan_b200_vgg.err | tail -1 > gpurun_out/r2/gan_b200_vgg_n1.json; head -c 350 gpurun_out/r2/gan_b200_vgg_n1.json; echo
timeout 600 python tools/gan_step.py --generator b200 --content reference --steps 10 2>/dev/null | tail -1 | head -c 350; echo
timeout 600 python -m pytest tests/test_vgg.py -q -m gpu 2>&1 | tail -2
