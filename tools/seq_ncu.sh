M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,lts__t_sector_hit_rate.pct
ncu --metrics $M --clock-control none -k regex:lsd_grow_kernel -s 1 -c 1 --csv --log-file gpurun_out/seq_ncu.csv python tools/prof_one.py 512 2 > /dev/null 2>&1
python tools/summarize_ncu_csv.py gpurun_out/seq_ncu.csv seq | grep lsd_grow
