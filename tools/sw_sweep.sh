# exactness sweep of the streaming small-batch grower (launches of <= 128 images): every output array against the oracle
set -x
python tools/parity_sweep.py 128 70000 rect 0 752 480 1
python tools/parity_sweep.py 128 71000 rect 0 752 480 2
python tools/parity_sweep.py 256 72000 rect 0 752 480 64
python tools/parity_sweep.py 96 73000 rect 0 752 480 8
python tools/parity_sweep.py 64 74000 curvy 0 752 480 1
python tools/parity_sweep.py 32 75000 curvy 0 641 479 3
python tools/parity_sweep.py 32 76000 rect 0 1280 720 1
python tools/parity_sweep.py 32 77000 rect 0 1241 376 4
