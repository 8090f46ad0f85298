for thr in 12 14 16 18 24; do echo thr $thr
for cfg in "dair_r50 64" "sgv3d_bsm_r50 16" "rope3d_r50 32"; do set -- $cfg
  SGV3D_BWD_THR=$thr timeout 120 python tools/time_kernels.py --shape $1 --batch $2 --pipeline tile --iters 30 2>&1 | sed -n 3p
done; done
