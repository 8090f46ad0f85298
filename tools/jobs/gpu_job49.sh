for bda in identity random none; do for p in tile block; do
timeout 120 python tools/time_kernels.py --shape dair_r50 --batch 64 --pipeline $p --iters 20 --bda $bda 2>&1 | sed -n 1p
done; done
timeout 120 python tools/time_kernels.py --shape sgv3d_bsm_r50 --batch 16 --pipeline tile --iters 20 --bda random 2>&1 | sed -n 1p
