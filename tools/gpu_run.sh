mkdir -p gpurun_out
for t in 1 2 3; do
  DSMIL_B200_TILES_PER_REC=$t timeout 300 python bench.py --no-extras --cpu-seconds 1 > gpurun_out/r2_bench_tpr$t.json 2> gpurun_out/r2_bench_tpr$t.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_tpr$t.json').read().strip().splitlines()[-1]); print('tiles/rec=$t', 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'])"
done
