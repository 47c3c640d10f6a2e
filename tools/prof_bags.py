"""One forward_bags call over the bench workload (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Weights, make_net
dev = torch.device("cuda", 0)
net = make_net(Weights(0), dev)
g = torch.Generator(device=dev).manual_seed(100)
bags = [torch.rand(10000, 512, generator=g, device=dev) for _ in range(16)]
with torch.no_grad():
    for _ in range(3):
        net.forward_bags(bags)
torch.cuda.synchronize()
