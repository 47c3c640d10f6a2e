mkdir -p gpurun_out
for v in epi2 epi4; do
  export DSMIL_B200_LIBPATH=$PWD/tools/variants/libdsmil_$v.so
  timeout 300 python bench.py --no-extras --cpu-seconds 1 > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_$v.json').read().strip().splitlines()[-1]); print('$v', 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'])"
done
