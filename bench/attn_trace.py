"""Per-step timeline of the heaviest forward-attention CTA (clock64 stamps of row thread 0)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relora_b200.ops import fused as F
C = F._C()
B, T, nh, hd = 24, 512, 16, 48
h = nh * hd
qkv = (torch.randn(B * T, 3 * h, device="cuda") * 0.5).bfloat16()
out = torch.empty(B * T, h, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, nh, T, device="cuda", dtype=torch.float32)
for _ in range(3): C.attention_fwd(qkv, out, lse, B, T, nh, hd, 1 / math.sqrt(hd))
tr = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
C.attention_set_trace(tr)
C.attention_fwd(qkv, out, lse, B, T, nh, hd, 1 / math.sqrt(hd)); torch.cuda.synchronize()
C.attention_set_trace(None)
t = tr.view(8, 64).cpu()
t0 = int(t[7, 0])
for jj in range(8):
    v = [int(t[i, jj]) - t0 for i in range(5)]
    print(f"step {jj}: start {v[0]} wait_s +{v[1]-v[0]} ld +{v[2]-v[1]} softmax +{v[3]-v[2]} storeP+arrive +{v[4]-v[3]}")
