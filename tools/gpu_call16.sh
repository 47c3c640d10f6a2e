mkdir -p gpurun_out
for f in 0 1 2 6; do
  echo "=== flags $f" 
  DSMIL_B200_PAIR=1 DSMIL_B200_PAIR_FLAGS=$f timeout 200 python tools/ptrace.py > gpurun_out/r2_ptrace_flags$f.txt 2>&1
  grep -E "steady state|TMA latency|mean" gpurun_out/r2_ptrace_flags$f.txt | cut -c1-260
  DSMIL_B200_PAIR=1 DSMIL_B200_PAIR_FLAGS=$f timeout 300 python bench.py --no-extras --cpu-seconds 1 > gpurun_out/r2_bench_pair_flags$f.json 2> /dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_pair_flags$f.json').read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], d['roofline']['per_kernel_ms'])"
done
