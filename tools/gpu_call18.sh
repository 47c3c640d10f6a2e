mkdir -p gpurun_out
DSMIL_B200_PAIR=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_pytest_gpu_10.log 2>&1; echo "pytest(pair) rc=$?" >> gpurun_out/r2_pytest_gpu_10.log
tail -3 gpurun_out/r2_pytest_gpu_10.log
for f in 0 2; do
  echo "=== flags $f" 
  DSMIL_B200_PAIR=1 DSMIL_B200_PAIR_FLAGS=$f timeout 200 python tools/ptrace.py > gpurun_out/r2_ptrace_v23_flags$f.txt 2>&1
  grep -E "steady state|mean" gpurun_out/r2_ptrace_v23_flags$f.txt | cut -c1-260
  DSMIL_B200_PAIR=1 DSMIL_B200_PAIR_FLAGS=$f timeout 300 python bench.py --no-extras --cpu-seconds 1 > gpurun_out/r2_bench_pair_v23_flags$f.json 2> /dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_pair_v23_flags$f.json').read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], d['roofline']['per_kernel_ms'])"
done
head -28 gpurun_out/r2_ptrace_v23_flags0.txt | cut -c1-120
