mkdir -p gpurun_out
timeout 120 tools/probe_pair 0 > gpurun_out/r2_probe_pair.txt 2>&1; echo "probe rc=$?" >> gpurun_out/r2_probe_pair.txt
timeout 120 tools/probe_pair 1 >> gpurun_out/r2_probe_pair.txt 2>&1; echo "probe(mode1) rc=$?" >> gpurun_out/r2_probe_pair.txt
cat gpurun_out/r2_probe_pair.txt
timeout 600 python -m pytest tests/test_zz_acceptance_gpu.py -q -p no:cacheprovider > gpurun_out/r2_pytest_accept.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_accept.log
tail -5 gpurun_out/r2_pytest_accept.log
