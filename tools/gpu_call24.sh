mkdir -p gpurun_out
for v in default a3w3 a2w4; do
  if [ $v = default ]; then unset DSMIL_B200_LIBPATH; else export DSMIL_B200_LIBPATH=$PWD/tools/variants/libdsmil_$v.so; fi
  timeout 300 python bench.py --no-extras --cpu-seconds 1 > gpurun_out/r2_bench_ring_$v.json 2> gpurun_out/r2_bench_ring_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_ring_$v.json').read().strip().splitlines()[-1]); print('$v', 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'])"
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "golden or bags or shapes" 2>&1 | tail -1
done
