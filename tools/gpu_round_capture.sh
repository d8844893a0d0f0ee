# Round capture on the GPU box: both bench arms (C2 default, C3), ncu launch lists and one --set full pass per kernel.
# Output (gpurun_out/) must stay under 64 MiB: one captured step per config.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err
timeout 600 python bench.py --config C3 --no-cpu-baseline > gpurun_out/bench_ours_C3.json 2> gpurun_out/bench_ours_C3.err
timeout 600 python bench.py --config C3 --impl reference > gpurun_out/bench_ref_C3.json 2> gpurun_out/bench_ref_C3.err
for c in C2 C3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1_launches_$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1_launches_$c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_|tile_sort|preprocess|scatter|tile_prefix|tile_scan" -s 99 -c 9 -f -o gpurun_out/r1_full_$c python bench.py --config $c --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1_full_$c.log 2>&1
done
ls -la gpurun_out/
