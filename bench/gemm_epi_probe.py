"""Epilogue diagnostics: time the K=128 (epilogue-bound) GEMM; run with RB_GEMM_DEBUG=0/1/2/4/7 in separate processes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relora_b200.ops import fused as F
M, N = 12288, 2304
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for K in (128, 768):
    x = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = {}
    for name, kw in (("bn256", dict(block_n=256, pair=0)), ("pair", dict(block_n=256, pair=1)), ("bn128", dict(block_n=128, pair=0))):
        for fl in (True, False):
            for _ in range(3): F.gemm(x, W, out, **kw)
            ts = []
            for _ in range(10):
                if fl: flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); F.gemm(x, W, out, **kw); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort(); res[name + ("_flush" if fl else "_warm")] = round(ts[5], 1)
    print("debug", os.environ.get("RB_GEMM_DEBUG", "0"), "K", K, res, flush=True)
