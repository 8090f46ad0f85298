for mk in 0xffffffff 63 1023; do
  SGV3D_REDUCE_ROWMASK=$mk timeout 300 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline tile 2>&1 | sed -n 2p | sed "s/^/rowmask=$mk /"
  SGV3D_REDUCE_ROWMASK=$mk timeout 300 python tools/time_kernels.py --shape sgv3d_bsm_r50 --batch 16 --pipeline tile 2>&1 | sed -n 2p | sed "s/^/rowmask=$mk /"
done
