for b in 1 4 8; do
  for p in tile block; do
    timeout 120 python tools/time_kernels.py --shape dair_r50 --batch $b --pipeline $p --iters 50 2>&1 | tail -1
  done
done
for b in 1 2; do for p in tile block; do timeout 120 python tools/time_kernels.py --shape sgv3d_bsm_r50 --batch $b --pipeline $p --iters 50 2>&1 | tr '\n' ' '; echo; done; done
