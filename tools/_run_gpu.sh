mkdir -p gpurun_out/r2
for tool in memcheck synccheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -4 gpurun_out/r2/sanitizer_$tool.log
done
