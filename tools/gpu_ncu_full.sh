# one `ncu --set full` capture (with source) of every hot kernel at the config in $1 (default C3); report -> gpurun_out/r2_full_$1.ncu-rep
c=${1:-C3}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"render_|tile_sort_dist|preprocess|scatter|tile_prefix|tile_scan" -s 32 -c 8 -f -o gpurun_out/r2_full_$c python bench.py --config $c --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_full_$c.log 2>&1
ls -la gpurun_out/r2_full_$c.ncu-rep; tail -3 gpurun_out/r2_full_$c.log
