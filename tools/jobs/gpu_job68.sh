mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ls_|camera_prep" -s 9 -c 9 -o gpurun_out/full_tile_final2 -f python tools/prof_step.py --batch 64 --steps 2 --backward > gpurun_out/ncu68.log 2>&1
tail -2 gpurun_out/ncu68.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches68.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/bench68_under_ncu.log 2>&1
tail -2 gpurun_out/launches68.csv | cut -c1-200
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench68_ref.json 2> gpurun_out/bench68_ref.err; cut -c1-300 gpurun_out/bench68_ref.json
