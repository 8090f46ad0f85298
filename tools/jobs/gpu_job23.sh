timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ls_plan_runs_fast" -s 1 -c 1 -o gpurun_out/plan_ev -f python tools/prof_step.py --batch 64 --steps 2 > gpurun_out/ncu23.log 2>&1
tail -3 gpurun_out/ncu23.log
