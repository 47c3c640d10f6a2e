mkdir -p gpurun_out
python - <<'PY'
import time, torch, sys
sys.path.insert(0,'.')
from bench import Weights, make_net, NBAG, D, C
from dsmil_wsi_b200.pipeline import HostBagPipeline
dev=torch.device('cuda',0); net=make_net(Weights(0),dev)
host=[torch.rand(NBAG,D).pin_memory() for _ in range(16)]
for cs,depth in ((1,2),(2,4),(2,6),(3,6)):
    pipe=HostBagPipeline(net,NBAG,D,C,depth=depth,copy_streams=cs)
    for _ in range(2): pipe.run(host)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(8): pipe.run(host)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/8
    print(f"copy_streams={cs} depth={depth}: {dt*1e3:.3f} ms/step  {16*NBAG/dt/1e6:.2f} M patches/s  H2D {16*NBAG*D*4/dt/1e9:.1f} GB/s")
PY
