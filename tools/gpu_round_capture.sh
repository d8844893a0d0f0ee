# Round-2 evidence capture on the GPU box (one GPU): both bench arms at the default config (C3), C2 / C4 / C5 for the tables,
# ncu launch lists and one `--set full` pass per kernel for C2 / C3 / C5.  gpurun_out/ must stay under 64 MiB.
set -x
mkdir -p gpurun_out
R=${ROUND:-r2}
timeout 600 python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${R}_bench_default_ref.json 2> gpurun_out/${R}_bench_default_ref.err
for c in C2 C4 C5; do
timeout 900 python bench.py --config $c --no-cpu-baseline > gpurun_out/${R}_bench_ours_$c.json 2> gpurun_out/${R}_bench_ours_$c.err
timeout 900 python bench.py --config $c --impl reference > gpurun_out/${R}_bench_ref_$c.json 2> gpurun_out/${R}_bench_ref_$c.err
done
for c in C3 C2 C5; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${R}_launches_$c.csv python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${R}_launches_$c.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"render_|tile_sort_dist|preprocess|scatter|tile_prefix|tile_scan" -s 32 -c 8 -f -o gpurun_out/${R}_full_$c python bench.py --config $c --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${R}_full_$c.log 2>&1
done
ls -la gpurun_out/ | tail -30
