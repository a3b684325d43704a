# ncu rows of the streaming grower, final code of round 2
set -x
ncu --set full --clock-control none --import-source on -k regex:lsd_grow_sw_kernel -s 1 -c 1 -o gpurun_out/r02f_grow_sw python tools/prof_one.py 64 2 > gpurun_out/r02f_ncu_grow_sw.log 2>&1
ncu -i gpurun_out/r02f_grow_sw.ncu-rep --page raw --csv > gpurun_out/r02f_grow_sw_raw.csv 2>/dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,lts__t_bytes.sum
ncu --metrics $M --clock-control none -s 31 -c 40 --csv --log-file gpurun_out/r02f_all_kernels_b1.csv python tools/prof_one.py 1 2 > gpurun_out/r02f_ncu_all_b1.log 2>&1
PLF_SW_FLAGS=4 python tools/latency_stages.py 1 2>&1 | grep "^sw " | tail -3 > gpurun_out/r02f_sw_counters.txt
