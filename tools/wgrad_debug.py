import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasfm_b200 import ops
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
for (E, Nout, Kout, erow) in [(16, 32, 32, 0), (16, 32, 32, 9), (16, 256, 64, 3)]:
    dy = torch.zeros(E, Nout, device=dev); x = torch.zeros(E, Kout, device=dev)
    dy[erow] = torch.arange(1, Nout + 1, device=dev).float()
    x[erow] = torch.arange(1, Kout + 1, device=dev).float() * 0.01
    got = ops.wgrad_tf32x3(dy, x); ref = dy.t() @ x
    print("case", E, Nout, Kout, erow, "max|got|", got.abs().max().item(), "err", (got - ref).abs().max().item())
    print("got[:6,:10]\n", got[:6, :10].cpu()); print("ref[:6,:10]\n", ref[:6, :10].cpu())
    nz = (got.abs() > 0).nonzero()
    print("nonzero count", nz.shape[0], "first", nz[:5].tolist())
dy = torch.randn(64, 32, device=dev); x = torch.randn(64, 32, device=dev)
got = ops.wgrad_tf32x3(dy, x); ref = dy.t() @ x
print("random: got[:3,:6]", got[:3, :6].cpu(), "\nref", ref[:3, :6].cpu())
