"""Run one GEMM configuration a few times (target for `ncu -k regex:gemm_kernel`)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relora_b200.ops import fused as F
ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=12288); ap.add_argument("--N", type=int, default=2304); ap.add_argument("--K", type=int, default=768)
ap.add_argument("--bn", type=int, default=128); ap.add_argument("--iters", type=int, default=6); ap.add_argument("--lora", type=int, default=0); ap.add_argument("--pair", type=int, default=0)
a = ap.parse_args()
x = torch.randn(a.M, a.K, device="cuda").to(torch.bfloat16); W = (torch.randn(a.N, a.K, device="cuda") * 0.02).to(torch.bfloat16)
out = torch.empty(a.M, a.N, device="cuda", dtype=torch.bfloat16)
u = torch.randn(a.M, 128, device="cuda").to(torch.bfloat16); B = (torch.randn(a.N, 128, device="cuda") * 0.02).to(torch.bfloat16)
for _ in range(a.iters):
    if a.lora: F.gemm(x, W, out, a2=u, b2=B, K2=128, block_n=a.bn, pair=a.pair)
    else: F.gemm(x, W, out, block_n=a.bn, pair=a.pair)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
