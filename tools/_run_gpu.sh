mkdir -p gpurun_out/r2
export SRGANFD_REFERENCE=$GRAFT_REPO_ROOT/baseline/_ref
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 "$@"; }
TR 8 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/r2/train_n8.err | tail -1 > gpurun_out/r2/train_n8.json; head -c 200 gpurun_out/r2/train_n8.json; echo
B200SR_DP_OVERLAP=1 TR 8 bench.py --gpus 8 --steps 20 --warmup 5 2> /dev/null | tail -1 > gpurun_out/r2/train_n8_overlap.json; head -c 200 gpurun_out/r2/train_n8_overlap.json; echo
for n in 2 4 8; do
  TR $n bench.py --workload c4 --gpus $n --steps 5 --warmup 3 2> gpurun_out/r2/c4_n$n.err | tail -1 > gpurun_out/r2/c4_n$n.json; head -c 260 gpurun_out/r2/c4_n$n.json; echo
done
TR 8 tools/gan_step.py --generator b200 --steps 10 2> gpurun_out/r2/gan_b200_n8.err | tail -1 > gpurun_out/r2/gan_b200_n8.json; head -c 330 gpurun_out/r2/gan_b200_n8.json; echo
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null > gpurun_out/r2/train_n1_samebox.json; head -c 200 gpurun_out/r2/train_n1_samebox.json; echo
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -2
