for b in 1 2 4 8; do
  for p in tile block; do
    timeout 120 python tools/time_kernels.py --shape dair_r50 --batch $b --pipeline $p --iters 50 2>&1 | head -2 | tr '\n' ' '; echo
  done
done
for p in tile block; do timeout 120 python tools/time_kernels.py --shape sgv3d_bsm_r50 --batch 1 --pipeline $p --iters 50 2>&1 | head -2 | tr '\n' ' '; echo; done
