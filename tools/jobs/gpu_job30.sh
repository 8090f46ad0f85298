for dbg in 0 1 2 3 8; do echo dbg $dbg; SGV3D_REDUCE_DBG=$dbg timeout 120 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline tile --iters 30 2>&1 | sed -n 2p; done
