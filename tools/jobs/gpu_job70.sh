for occ in 2 3 4; do echo occ $occ; for cfg in "dair_r50 64" "rope3d_r50 32" "sgv3d_bsm_r50 16"; do set -- $cfg
SGV3D_BWD_OCC=$occ timeout 120 python tools/time_kernels.py --shape $1 --batch $2 --pipeline tile --iters 30 2>&1 | sed -n 3p | cut -c1-110; done; done
