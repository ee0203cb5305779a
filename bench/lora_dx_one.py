"""Run the fused LoRA input-gradient kernel a few times (target for `ncu -k regex:lora_dx_fused2` or `regex:lora_dx_pair`)."""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relora_b200.ops import fused as F
ap = argparse.ArgumentParser(); ap.add_argument("--M", type=int, default=12288); ap.add_argument("--K", type=int, default=2560)
ap.add_argument("--Ng", type=int, default=768); ap.add_argument("--G", type=int, default=1); ap.add_argument("--iters", type=int, default=6)
a = ap.parse_args(); C = F._C(); r = 128
dy = torch.randn(a.M, a.G * a.Ng, device="cuda").bfloat16(); W = (torch.randn(a.G * a.Ng, a.K, device="cuda") * 0.02).bfloat16()
du = torch.randn(a.M, a.G * r, device="cuda").bfloat16(); A = (torch.randn(a.G * r, a.K, device="cuda") * 0.02).bfloat16()
out = torch.empty(a.M, a.K, device="cuda", dtype=torch.bfloat16); seed = torch.tensor([7], dtype=torch.int32, device="cuda")
for _ in range(a.iters): C.lora_dx(dy, W, du, A, out, seed, list(range(1, a.G + 1)), 0.1)
torch.cuda.synchronize(); print("ok", float(out.float().abs().mean()))
