"""Run the decoder forward+backward a few times at config-3 size (for ncu captures of the MLP kernels)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianavatar_b200.network import POP_no_unet
torch.manual_seed(0)
net = POP_no_unet(c_geom=64, hsize=128).cuda()
geo = (torch.randn(1, 64, 128, 128) * 0.01).cuda().requires_grad_(True)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for _ in range(n):
    dec = net.forward_packed(geo, S, 2)
    dec.backward(torch.randn_like(dec))
torch.cuda.synchronize()
print("ok")
