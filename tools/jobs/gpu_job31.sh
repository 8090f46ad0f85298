for dbg in 0 16; do echo dbg $dbg; SGV3D_REDUCE_DBG=$dbg timeout 120 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline tile --iters 30 2>&1 | sed -n 2p; 
SGV3D_REDUCE_DBG=$dbg timeout 120 python tools/time_kernels.py --shape sgv3d_bsm_r50 --batch 16 --pipeline tile --iters 30 2>&1 | sed -n 2p; done
