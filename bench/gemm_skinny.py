"""Narrow-output / short-M GEMM sites of the LoRA branch vs cuBLAS, plus the CTA-0 timeline of each shape (kernel entry -> first
load -> first / last full barrier -> epilogue -> exit, and the phases of the first output slab).

    python bench/gemm_skinny.py [--out gpurun_out/x.json]

Inputs rotate over 8 buffer sets (> L2 capacity in total) so every call streams its operands from HBM, as in the training step.
"""
import argparse, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relora_b200.ops import fused as F
ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); ap.add_argument("--iters", type=int, default=40)
a = ap.parse_args()
C = F._C()
dev = "cuda"
NB = 8

def timed(fn):
    for i in range(NB): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters): fn(i % NB)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters * 1e3

def trace(fn):
    tr = torch.zeros(16 * 512, dtype=torch.int64, device=dev); tr[14 * 512] = 2 ** 62
    C.gemm_set_trace(tr); fn(0); torch.cuda.synchronize(); C.gemm_set_trace(None)
    t = tr.view(16, 512).cpu()
    ent, setup, ex = int(t[12, 0]), int(t[12, 1]), int(t[12, 2])
    loads = [int(x) - ent for x in t[0] if int(x)]; full = [int(x) - ent for x in t[1] if int(x)]
    es = [int(x) - ent for x in t[4] if int(x)]; ee = [int(x) - ent for x in t[5] if int(x)]
    return {"cta0_cycles": ex - ent, "setup": setup - ent, "first_load": loads[:1], "last_load": loads[-1:], "first_full": full[:1], "last_full": full[-1:],
            "epi_start": es[:2], "epi_end": ee[:2], "slab0_phases": [int(t[r, 0]) - int(t[4, 0]) for r in range(6, 12)], "n_tiles_cta0": len(ee), "cta0_ns": int(t[13, 1] - t[13, 0]), "grid_span_ns": int(t[15, 0] - t[14, 0])}

rows = []
# (name, M, N, K, a_mn, b_mn, acc32)
shapes = [("u = xd*A^T   (o)", 12288, 128, 768, False, False, False), ("du = dy*B   (o,down)", 12288, 128, 768, False, True, False),
          ("u = hd*A^T   (down)", 12288, 128, 2560, False, False, False), ("du (gate/up, one group)", 12288, 128, 2560, False, True, False),
          ("u (qkv, N=384)", 12288, 384, 768, False, False, False), ("u (gate/up, N=256)", 12288, 256, 768, False, False, False),
          ("dB = dy^T*u (o)", 768, 128, 12288, True, True, True), ("dA = du^T*xd (o)", 128, 768, 12288, True, True, True),
          ("dB (down)", 768, 128, 12288, True, True, True), ("dA (down)", 128, 2560, 12288, True, True, True)]
for name, M, N, K, amn, bmn, acc in shapes:
    As = [(torch.randn((K, M) if amn else (M, K), device=dev)).bfloat16() for _ in range(NB)]
    Bs = [(torch.randn((K, N) if bmn else (N, K), device=dev) * 0.05).bfloat16() for _ in range(NB)]
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if acc else torch.bfloat16)
    rec = {"site": name, "M": M, "N": N, "K": K, "a_mn": amn, "b_mn": bmn, "acc32": acc}
    def ref(i):
        x = As[i].t() if amn else As[i]; w = Bs[i] if bmn else Bs[i].t()
        return torch.matmul(x, w)
    rec["cublas_us"] = timed(ref)
    for bn in (128,):
        def run(i, bn=bn):
            F.gemm(As[i], Bs[i], out, M=M, N=N, K1=K, a1_mn=amn, b1_mn=bmn, accumulate=acc, out_dtype=out.dtype, block_n=bn, split_k=0 if acc else 1, pair=0)
        try:
            rec[f"bn{bn}_us"] = timed(run)
            if not acc:
                out.zero_(); run(0); err = (out.float() - ref(0).float()).abs().max().item(); rec[f"bn{bn}_maxerr"] = err
            else:
                out.zero_(); run(0); err = (out - ref(0).float()).abs().max().item() / (ref(0).float().abs().max().item() + 1e-9); rec[f"bn{bn}_relerr"] = err
            rec[f"bn{bn}_trace"] = trace(run)
        except Exception as e:  # noqa: BLE001
            rec[f"bn{bn}_error"] = repr(e)[:200]
    byts = (M * K + N * K) * 2 + M * N * (4 if acc else 2)
    rec["hbm_floor_us"] = byts / 6571.9e9 * 1e6
    rows.append(rec)
    print(json.dumps(rec), flush=True)
if a.out:
    os.makedirs(os.path.dirname(a.out), exist_ok=True); json.dump(rows, open(a.out, "w"), indent=1)
