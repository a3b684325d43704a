# Rebuilds the tracked round-2 profile files from the scratch outputs of tools/r02_capture.sh (run here, after gpurun).
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/r02_parity_sweep.txt profiles/r02_parity_sweep.txt
cp gpurun_out/r02_launches.csv profiles/r02_launch_list_bench.csv
python tools/summarize_launches.py gpurun_out/r02_launches.csv "Round 2 (end) — ncu launch list of \`python bench.py --steps 2 --warmup 3 --no-cpu-baseline\` (first 400 launches: 8 contexts x 512 pairs, B200; kernels inside the replayed CUDA graphs are profiled as graph nodes)" > profiles/r02_launch_list_bench.md
python tools/summarize_ncu_csv.py gpurun_out/r02_all_kernels.csv "counters" > /tmp/r02_all_table.md
python tools/summarize_ncu_csv.py gpurun_out/r02_c3_dominant.csv "c3" > /tmp/r02_c3_table.md
python tools/summarize_ncu_csv.py gpurun_out/r02_c4_dominant.csv "c4" > /tmp/r02_c4_table.md
echo done
