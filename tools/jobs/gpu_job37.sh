for r in 0 1 2; do echo role $r; for cfg in "dair_r50 64" "sgv3d_bsm_r50 16"; do set -- $cfg
SGV3D_PREP_ROLE=$r timeout 120 python tools/time_kernels.py --shape $1 --batch $2 --pipeline tile --iters 30 2>&1 | sed -n 2p; done; done
