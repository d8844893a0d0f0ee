# final-state evidence for the default config only (budget-friendly subset of gpu_round_capture.sh)
mkdir -p gpurun_out
R=${ROUND:-r2}
timeout 600 python bench.py > gpurun_out/${R}_final_default.json 2> gpurun_out/${R}_final_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${R}_final_default_ref.json 2> gpurun_out/${R}_final_default_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${R}_launches_C3.csv python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${R}_launches_C3.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"render_|tile_sort_dist|preprocess|scatter|tile_prefix|tile_scan" -s 32 -c 8 -f -o gpurun_out/${R}_full_C3 python bench.py --config C3 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${R}_full_C3.log 2>&1
