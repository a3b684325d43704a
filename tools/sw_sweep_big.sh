# larger exactness sweep of the streaming grower: 2080 more pairs at odd batch sizes (every output array against the oracle)
set -x
python tools/parity_sweep.py 512 90000 rect 0 752 480 1
python tools/parity_sweep.py 400 91000 rect 0 752 480 5
python tools/parity_sweep.py 512 92000 rect 0 752 480 16
python tools/parity_sweep.py 296 93000 rect 0 752 480 37
python tools/parity_sweep.py 148 94000 rect 0 752 480 148
python tools/parity_sweep.py 128 95000 curvy 0 752 480 2
python tools/parity_sweep.py 84 96000 curvy 0 752 480 7
