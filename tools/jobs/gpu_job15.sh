python tools/time_op.py --batch 4; python tools/time_op.py --batch 16
