"""Pipeline timeline of CTA 0 of one tcgen05 GEMM launch (clock64 stamps written by the kernel's own roles).

    python bench/gemm_trace.py [--M 12288 --N 2304 --K 768 --bn 256 --pair 0]
"""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relora_b200.ops import fused as F
ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=12288); ap.add_argument("--N", type=int, default=2304); ap.add_argument("--K", type=int, default=768)
ap.add_argument("--bn", type=int, default=256); ap.add_argument("--pair", type=int, default=0)
a = ap.parse_args()
C = F._C()
x = torch.randn(a.M, a.K, device="cuda").bfloat16(); W = (torch.randn(a.N, a.K, device="cuda") * 0.02).bfloat16()
out = torch.empty(a.M, a.N, device="cuda", dtype=torch.bfloat16)
for _ in range(3): F.gemm(x, W, out, block_n=a.bn, pair=a.pair)
tr = torch.zeros(16 * 512, dtype=torch.int64, device="cuda")
C.gemm_set_trace(tr)
F.gemm(x, W, out, block_n=a.bn, pair=a.pair)
torch.cuda.synchronize()
C.gemm_set_trace(None)
t = tr.view(16, 512).cpu()
t0 = int(t[0, 0])
names = ["load_issue", "mma_full", "acc_free", "tile_commit", "epi_start", "epi_end"]
kb = (a.K + 63) // 64
def rel(v): return [int(x) - t0 for x in v if int(x) != 0]
P, Mf, Af, Tc, Es, Ee = (rel(t[i]) for i in range(6))
print(f"K={a.K} k-blocks/tile={kb}; cycles relative to the first load issue (CTA 0)")
for i in range(min(6, len(Tc))):
    ps = P[i * kb:(i + 1) * kb]; ms = Mf[i * kb:(i + 1) * kb]
    print(f"tile {i}: loads {ps[:3]}..{ps[-1:]}  mma_full {ms[:3]}..{ms[-1:]}  acc_free {Af[i]}  commit_issued {Tc[i]}  epi {Es[i]}->{Ee[i]} ({Ee[i]-Es[i]})")
S = [rel(t[i]) for i in range(6, 12)]
for i in range(min(6, len(Tc))):
    try:
        print(f"tile {i} slab0: epi_start {Es[i]} ld_done +{S[0][i]-Es[i]} math +{S[1][i]-S[0][i]} sts +{S[2][i]-S[1][i]} fence +{S[3][i]-S[2][i]} bar +{S[4][i]-S[3][i]} store+wait_read +{S[5][i]-S[4][i]}")
    except IndexError:
        pass
if len(Tc) > 2:
    print("steady state per tile (cycles):", [Ee[i + 1] - Ee[i] for i in range(len(Ee) - 1)])
    print("load->full latency (cycles):", [m - p for p, m in zip(P[:12], Mf[:12])])
