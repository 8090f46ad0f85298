for dbg in 0 1 2 4 8 7 15 16; do echo dbg $dbg; SGV3D_BWD_DBG=$dbg timeout 120 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline tile --iters 30 2>&1 | sed -n 3p; done
