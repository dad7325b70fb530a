mkdir -p gpurun_out/r2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_final_a.json 2> gpurun_out/r2/bench_final_a.err; head -c 300 gpurun_out/r2/bench_final_a.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/bench_ref_a.json 2>/dev/null; head -c 300 gpurun_out/r2/bench_ref_a.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_final.csv python tools/profile_step.py > gpurun_out/r2/ncu_lf.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:conv3x3_chain --launch-skip 4 --launch-count 2 -o gpurun_out/r2/chain_final -f python tools/profile_step.py > gpurun_out/r2/ncu_cf.log 2>&1
ncu --set full --clock-control none -k regex:wgrad3x3 --launch-skip 170 --launch-count 2 -o gpurun_out/r2/wgrad_final -f python tools/profile_step.py > gpurun_out/r2/ncu_wf.log 2>&1
tail -2 gpurun_out/r2/ncu_cf.log gpurun_out/r2/ncu_wf.log
