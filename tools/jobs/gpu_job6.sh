mkdir -p gpurun_out
for occ in 2 3; do
  SGV3D_BWD_OCC=$occ timeout 300 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline tile 2>&1 | tail -1 | sed "s/^/occ=$occ /"
  SGV3D_BWD_OCC=$occ timeout 300 python tools/time_kernels.py --shape sgv3d_bsm_r50 --batch 16 --pipeline tile 2>&1 | tail -1 | sed "s/^/occ=$occ /"
done > gpurun_out/t6.log 2>&1
timeout 300 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline tile 2>&1 | head -1 >> gpurun_out/t6.log
timeout 300 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline block 2>&1 >> gpurun_out/t6.log
cat gpurun_out/t6.log
