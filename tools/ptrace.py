"""CTA-0 timeline of the pair kernel k_fwd_pair (clock64 stamps via dsmil_debug_set_trace; buffer 4 x 8 x 128 int64).
Usage: python tools/ptrace.py [bags] [rows]   -- prints per-role waiting / working ns for the first boxes and tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import Weights, make_net
from dsmil_wsi_b200 import _lib

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
dev = torch.device("cuda", 0)
net = make_net(Weights(0), dev)
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(100)
bags = [torch.rand(N, 512, generator=g, device=dev) for _ in range(nb)]
with torch.no_grad():
    for _ in range(3):
        net.forward_bags(bags)
    buf = torch.zeros(8 * 8 * 128 + 3 * 8 * 32, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    lib.dsmil_debug_set_trace(buf.data_ptr())
    net.forward_bags(bags)
    torch.cuda.synchronize()
    lib.dsmil_debug_set_trace(None)
full = buf.cpu().numpy().astype(np.int64)
t = full[:8192].reshape(8, 8, 128)
allw = full[8192:].reshape(3, 8, 2, 16)
nz = t[t > 0]
t0 = int(nz.min())
rel = lambda v: int(v - t0) if v else None
conv, mma, epi, prod = t[0], t[1], t[2], t[3]
np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "ptrace_raw.npy"), full)
mvalid = int((conv[4] > 0).sum())
print(f"converter group 0 / quadrant 0: {mvalid} boxes traced (every 4th box of the CTA)")
print(" m  wait_start  XS_FULL(+wait)  converted(+work)  XT_EMPTY(+wait)  stored(+work)   [ns]")
for m in range(min(mvalid, 24)):
    a0, a1, a2, a3, a4 = (int(conv[e, m]) for e in range(5))
    print(f"{m:2d}  {rel(a0):9d}  +{a1 - a0:6d}  +{a2 - a1:6d}  +{a3 - a2:6d}  +{a4 - a3:6d}")
if mvalid > 8:
    d = np.diff(conv[0, 4:mvalid])
    w_xs = (conv[1, 4:mvalid] - conv[0, 4:mvalid]); w_xt = (conv[3, 4:mvalid] - conv[2, 4:mvalid])
    wk = (conv[2, 4:mvalid] - conv[1, 4:mvalid]); st = (conv[4, 4:mvalid] - conv[3, 4:mvalid])
    print(f"steady state per box of one group: period {d.mean():.0f} ns = XS_FULL wait {w_xs.mean():.0f} + convert {wk.mean():.0f} "
          f"+ XT_EMPTY wait {w_xt.mean():.0f} + tcgen05.st/arrive {st.mean():.0f}")
nvalid = int((mma[1] > 0).sum())
print(f"MMA issuer: {nvalid} boxes traced;  XT_FULL-ready -> committed, and gap to the next box")
for n in range(min(nvalid, 20)):
    print(f"{n:3d}  ready {rel(int(mma[0, n])):9d}  issue {int(mma[1, n] - mma[0, n]):5d}  next-ready +{int(mma[0, n + 1] - mma[1, n]) if n + 1 < nvalid else 0:6d}")
if nvalid > 40:
    per = np.diff(mma[0, 16:nvalid]); print(f"steady state: {per.mean():.0f} ns per box at the MMA issuer ({16 * per.mean():.0f} per tile)")
jv = int((epi[3] > 0).sum())
print("epilogue warp 0: per tile  start-wait  H1_FULL(+wait)  A2 written(+work)  Q_FULL(+wait)  Q stored(+work)")
for j in range(min(jv, 8)):
    e4, e0, e1, e2, e3 = (int(epi[e, j]) for e in (4, 0, 1, 2, 3))
    print(f"{j:2d}  {rel(e4):9d}  +{e0 - e4:6d}  +{e1 - e0:6d}  +{e2 - e1:6d}  +{e3 - e2:6d}")
print("MMA layer 2: per tile  reached, A2_FULL wait; ACC_EMPTY granted")
for j in range(min(jv, 8)):
    print(f"{j:2d}  l2 reached {rel(int(mma[2, j]))}  waited {int(mma[3, j] - mma[2, j]) if mma[3, j] else None}  acc_empty {rel(int(mma[4, j]))}")
pv = int((prod[0] > 0).sum())
if pv > 8:
    print(f"producer: slot granted every {np.diff(prod[0, 3:pv]).mean():.0f} ns on average over {pv} boxes")

print("\n== cross-CTA view (globaltimer ns): when each traced converter warp ARRIVED on XT_FULL for its box m (box n = 2m for")
print("   group 0 warps [cw 0], n = 2m+1 for group 1 warps [cw 15]) vs when the MMA issuer saw the box ready ==")
print(" n   cta0.cw0  cta1.cw0 | cta0.cw15 cta1.cw15 |  mma ready  mma committed")
for n in range(0, 40):
    m = n // 4
    a0 = a1 = b0 = b1 = None
    if n % 4 == 0:
        a0, a1 = rel(int(t[0, 4, m])), rel(int(t[4, 4, m]))
    elif n % 4 == 3:
        b0, b1 = rel(int(t[5, 4, m])), rel(int(t[6, 4, m]))
    print(f"{n:2d}  {str(a0):>9} {str(a1):>9} | {str(b0):>9} {str(b1):>9} | {rel(int(mma[0, n])):9d} {rel(int(mma[1, n])):9d}")
print("\nXT_EMPTY seen (event 3) by cta0.cw0 / cta1.cw0 for box m vs MMA commit of box n-4:")
for m in range(1, 12):
    n = 4 * m
    print(f"m={m:2d} n={n:2d}  cta0 {rel(int(t[0, 3, m]))}  cta1 {rel(int(t[4, 3, m]))}  commit(n-4) {rel(int(mma[1, n - 4]))}")
e7 = t[7]
print("\nepilogue warp 0 of CTA 1: per tile  H1_FULL seen, A2 arrived, Q_FULL seen, Q stored")
for j in range(6):
    print(j, [rel(int(e7[e, j])) for e in (0, 1, 2, 3)], " cta0:", [rel(int(epi[e, j])) for e in (0, 1, 2, 3)])

print("\n== every converter warp, boxes 40..47 (ns since t0): rows = cta x warp (cw = g*8 + h*4 + q) ==")
for n in range(40, 48):
    g = n % 4
    print(f"box {n} (group {g}): mma ready {rel(int(mma[0, n]))}  committed {rel(int(mma[1, n]))}")
    for ev, name in enumerate(("XS_FULL seen ", "XT_EMPTY seen", "XT_FULL arrvd")):
        for c in range(2):
            vals = [rel(int(allw[ev, n - 40, c, cw])) for cw in range(g * 4, g * 4 + 4)]
            print(f"   {name} cta{c}: {vals}")

print("\nTMA latency estimate for group-0 boxes (slot granted to the producer -> XS_FULL seen by a waiting converter):")
lat = []
for m in range(2, min(mvalid, 30)):
    n = 4 * m
    if prod[0, n] and conv[1, m]:
        lat.append(int(conv[1, m] - prod[0, n]))
print("  ", lat[:28], " mean", (sum(lat) / max(len(lat), 1)))
