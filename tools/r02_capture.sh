# Round-2 evidence run on the GPU box: exactness sweeps, ncu launch list of the bench command, ncu captures of the dominant
# kernel of every benchmarked config and a counter row per kernel of one pass.  Outputs stay small (gpurun_out <= 64 MiB).
set -x
mkdir -p gpurun_out
( python tools/parity_sweep.py 1536 30000 rect 0 752 480 256; python tools/parity_sweep.py 512 40000 rect 0 752 480 64; python tools/parity_sweep.py 256 50000 rect 1 752 480 64; python tools/parity_sweep.py 128 60000 curvy 0 752 480 32; python tools/parity_sweep.py 64 61000 curvy 1 641 479 16; python tools/parity_sweep.py 64 62000 rect 0 1280 720 64; python tools/parity_sweep.py 64 63000 rect 0 1241 376 4 ) > gpurun_out/r02_parity_sweep.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsd_grow_kernel -s 1 -c 1 -o gpurun_out/r02_grow python tools/prof_one.py 512 2 > gpurun_out/r02_ncu_grow.log 2>&1
ncu -i gpurun_out/r02_grow.ncu-rep --page raw --csv > gpurun_out/r02_grow_raw.csv 2>/dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,lts__t_bytes.sum
ncu --metrics $M --clock-control none -s 31 -c 40 --csv --log-file gpurun_out/r02_all_kernels.csv python tools/prof_one.py 512 2 > gpurun_out/r02_ncu_all.log 2>&1
ncu --metrics $M --clock-control none -k regex:"fast_score|fast_cells" -s 2 -c 2 --csv --log-file gpurun_out/r02_c3_dominant.csv python tools/prof_one.py 512 2 752 480 2000 0 1 0 > gpurun_out/r02_ncu_c3.log 2>&1
ncu --metrics $M --clock-control none -k regex:lsd_grow_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_c4_dominant.csv python tools/prof_one.py 128 2 1280 720 1200 500 0 1 > gpurun_out/r02_ncu_c4.log 2>&1
rm -f gpurun_out/r02_all.ncu-rep
ls -la gpurun_out | tail -15
tail -3 gpurun_out/r02_parity_sweep.txt
