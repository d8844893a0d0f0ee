# compute-sanitizer on config C1 forward + backward (and a small quantised / masked case): memcheck, racecheck, synccheck.
# GPU box:  gpurun -- 'bash tools/gpu_sanitizer.sh'   -> gpurun_out/sanitizer_c1_<tool>.txt ; summarise into profiles/sanitizer_c1.txt
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=gsb --print-limit 20 python tools/sanitize_c1.py > gpurun_out/sanitizer_c1_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|C1 R|quant\+mask R" gpurun_out/sanitizer_c1_$tool.txt
done
