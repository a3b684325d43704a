set -x
mkdir -p gpurun_out
( python tools/parity_sweep.py 2048 30000 rect 0 752 480 256; python tools/parity_sweep.py 512 40000 rect 0 752 480 64; python tools/parity_sweep.py 256 50000 rect 1 752 480 64; python tools/parity_sweep.py 128 60000 curvy 0 752 480 32; python tools/parity_sweep.py 64 61000 curvy 1 641 479 16; python tools/parity_sweep.py 64 62000 rect 0 1280 720 64; python tools/parity_sweep.py 64 63000 rect 0 1241 376 4 ) > gpurun_out/r02_parity_sweep.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsd_grow_kernel -s 1 -c 1 -o gpurun_out/r02_grow python tools/prof_one.py 512 2 > gpurun_out/r02_ncu_grow.log 2>&1
ncu --set full --clock-control none -s 30 -c 30 -o gpurun_out/r02_all python tools/prof_one.py 512 2 > gpurun_out/r02_ncu_all.log 2>&1
tail -3 gpurun_out/r02_parity_sweep.txt
