mkdir -p gpurun_out
timeout 120 tools/probe_pair 0 > gpurun_out/r2_probe_pair.txt 2>&1; echo "probe rc=$?" >> gpurun_out/r2_probe_pair.txt
timeout 120 tools/probe_pair 1 >> gpurun_out/r2_probe_pair.txt 2>&1; echo "probe(mode1) rc=$?" >> gpurun_out/r2_probe_pair.txt
cat gpurun_out/r2_probe_pair.txt
