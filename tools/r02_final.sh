# End-of-round-2 evidence run on the GPU box (after the streaming small-batch grower and the forked line path): bench lines of
# every config, the reference arm, ncu launch list of the bench command, ncu captures of both growers, a counter row per
# kernel of one pass, compute-sanitizer over every entry point.  Outputs stay small (gpurun_out <= 64 MiB).
set -x
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_ref.err
python bench.py --steps 8 --warmup 3 > gpurun_out/r02f_bench_c2.json 2> gpurun_out/r02f_bench_c2.err
for c in c3 c4 c5 c2_batch64; do python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_$c.json 2> gpurun_out/r02f_bench_$c.err; done
python bench.py --rectify --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_rectify.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsd_grow_kernel -s 1 -c 1 -o gpurun_out/r02f_grow python tools/prof_one.py 512 2 > gpurun_out/r02f_ncu_grow.log 2>&1
ncu -i gpurun_out/r02f_grow.ncu-rep --page raw --csv > gpurun_out/r02f_grow_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:lsd_grow_sw_kernel -s 1 -c 1 -o gpurun_out/r02f_grow_sw python tools/prof_one.py 64 2 > gpurun_out/r02f_ncu_grow_sw.log 2>&1
ncu -i gpurun_out/r02f_grow_sw.ncu-rep --page raw --csv > gpurun_out/r02f_grow_sw_raw.csv 2>/dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,lts__t_bytes.sum
ncu --metrics $M --clock-control none -s 31 -c 40 --csv --log-file gpurun_out/r02f_all_kernels.csv python tools/prof_one.py 512 2 > gpurun_out/r02f_ncu_all.log 2>&1
ncu --metrics $M --clock-control none -s 31 -c 40 --csv --log-file gpurun_out/r02f_all_kernels_b1.csv python tools/prof_one.py 1 2 > gpurun_out/r02f_ncu_all_b1.log 2>&1
python tools/latency_stages.py 1 2 8 64 128 148 > gpurun_out/r02f_latency_stages.log 2>&1
timeout 420 compute-sanitizer --tool memcheck python tools/sanity_all.py > gpurun_out/r02f_san_mem.log 2>&1; tail -3 gpurun_out/r02f_san_mem.log
timeout 420 compute-sanitizer --tool racecheck python tools/sanity_all.py > gpurun_out/r02f_san_race.log 2>&1; tail -3 gpurun_out/r02f_san_race.log
rm -f gpurun_out/r02f_grow.ncu-rep.tmp
ls -la gpurun_out | grep r02f
