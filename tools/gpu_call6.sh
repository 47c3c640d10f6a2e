mkdir -p gpurun_out
: > gpurun_out/r2_probe_pair.txt
for m in 0 1; do
  timeout 120 tools/probe_pair $m >> gpurun_out/r2_probe_pair.txt 2>&1; echo "probe mode $m rc=$?" >> gpurun_out/r2_probe_pair.txt
done
cat gpurun_out/r2_probe_pair.txt
