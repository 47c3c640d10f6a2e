import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import build_net
from oracle import dsmil_oracle as orc
p = orc.random_params(512, 2, 41, scale=2.0)
net = build_net(p).eval()
sizes = [1, 127, 128, 129, 1000]
Xs = [orc.synthetic_bag(n, 512, 600 + i, "uniform" if i % 2 else "normal") for i, n in enumerate(sizes)]
xs = [torch.from_numpy(x).cuda() for x in Xs]
with torch.no_grad():
    o1 = net.forward_bags(xs); o2 = net.forward_bags(xs)
    s1 = [net(x) for x in xs]; s2 = [net(x) for x in xs]
names = ["classes","pred","A","B"]
for i in range(len(sizes)):
    print(sizes[i], "batch-vs-batch", [bool(torch.equal(a,b)) for a,b in zip(o1[i],o2[i])],
          "single-vs-single", [bool(torch.equal(a,b)) for a,b in zip(s1[i],s2[i])],
          "batch-vs-single", [bool(torch.equal(a,b)) for a,b in zip(o1[i],s1[i])],
          [float((a-b).abs().max()) for a,b in zip(o1[i],s1[i])])
