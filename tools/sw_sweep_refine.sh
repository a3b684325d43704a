# exactness of the streaming grower with lsd_refine = 1 (regions that need refining are grown and refined by the committing warp)
set -x
python tools/parity_sweep.py 96 50000 rect 1 752 480 1
python tools/parity_sweep.py 128 51000 rect 1 752 480 8
python tools/parity_sweep.py 128 52000 rect 1 752 480 64
python tools/parity_sweep.py 48 53000 curvy 1 641 479 3
python tools/parity_sweep.py 32 54000 rect 1 1280 720 2
