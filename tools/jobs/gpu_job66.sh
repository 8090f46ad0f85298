for thr in 0 6 9 11 14 18; do echo thr $thr
for cfg in "dair_r50 64" "rope3d_r50 32" "dair_r50 8"; do set -- $cfg
  SGV3D_BWD_THR=$thr timeout 120 python tools/time_kernels.py --shape $1 --batch $2 --pipeline tile --iters 30 2>&1 | sed -n 3p | cut -c1-100
done; done
