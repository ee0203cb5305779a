"""Run the two-kernel form of the LoRA input gradient (frozen-path product supplied, `lora_dx ... Kb0 (+base)`) a few times
(target for `ncu -k regex:lora_dx_base`).  Default shape: llama_250m qkv group (G = 3, N = 768)."""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relora_b200.ops import fused as F
ap = argparse.ArgumentParser(); ap.add_argument("--M", type=int, default=12288); ap.add_argument("--N", type=int, default=768)
ap.add_argument("--G", type=int, default=3); ap.add_argument("--iters", type=int, default=6)
a = ap.parse_args(); C = F._C(); r = 128
du = torch.randn(a.M, a.G * r, device="cuda").bfloat16(); A = (torch.randn(a.G * r, a.N, device="cuda") * 0.02).bfloat16()
base = torch.randn(a.M, a.N, device="cuda").bfloat16(); out = torch.empty(a.M, a.N, device="cuda", dtype=torch.bfloat16)
seed = torch.tensor([7], dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): C.lora_dx(None, None, du, A, out, seed, list(range(1, a.G + 1)), 0.1, base)
torch.cuda.synchronize(); e0.record()
for _ in range(a.iters): C.lora_dx(None, None, du, A, out, seed, list(range(1, a.G + 1)), 0.1, base)
e1.record(); torch.cuda.synchronize(); print("ok", float(out.float().abs().mean()), "us/call", e0.elapsed_time(e1) * 1e3 / a.iters)
