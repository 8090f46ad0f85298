timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
(timeout 1500 bash tools/sanitize.sh > gpurun_out/sanitize27.log 2>&1; tail -15 gpurun_out/sanitize27.log)
