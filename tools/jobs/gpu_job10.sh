mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"ls_|camera_prep" -s 9 -c 9 -o gpurun_out/full_tile -f python tools/prof_step.py --batch 64 --steps 2 --backward > gpurun_out/ncu10.log 2>&1
tail -2 gpurun_out/ncu10.log
