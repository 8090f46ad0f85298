(timeout 1500 bash tools/sanitize.sh > gpurun_out/sanitize29.log 2>&1; grep -v "Host Frame" gpurun_out/sanitize29.log | tail -12 | cut -c1-200)
