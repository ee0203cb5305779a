"""Busy / idle analysis of one captured micro-step replay: kineto kernel intervals of a CUDA-graph replay, union of the
intervals per stream and overall (how much of the span has no kernel running, how much runs two streams at once)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); os.environ.setdefault("RELORA_B200_NO_WANDB", "1")
from relora_b200.engine.api import TrainingEngine
from relora_b200.parallel.dist import DistInfo
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--model", default="llama_250m"); ap.add_argument("--batch", type=int, default=24)
a = ap.parse_args()
info = DistInfo(0, 0, 1, torch.device("cuda", 0), "nccl"); torch.cuda.set_device(0)
eng = TrainingEngine.build(info, model_config=os.path.join(ROOT, "configs", f"{a.model}.json"), batch_size=a.batch, gradient_accumulation=2,
    total_batch_size=2 * a.batch, max_length=512, use_peft=True, lora_r=128, relora=5000, cycle_length=5000, scheduler="cosine_restarts",
    warmup_steps=500, restart_warmup_steps=100, lr=1e-3, num_training_steps=20000, dtype="bfloat16", device="cuda", cuda_graphs=True, engine="fused")
ids = torch.randint(0, 32000, (2, a.batch, 512), device="cuda")
for _ in range(3): eng.train_step_device(ids)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.train_step_device(ids); torch.cuda.synchronize()
ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort()
t0, t1 = ev[0][0], max(e[1] for e in ev)
def union(iv):
    tot, cur_s, cur_e = 0.0, None, None
    for s, e in sorted(iv):
        if cur_e is None or s > cur_e:
            if cur_e is not None: tot += cur_e - cur_s
            cur_s, cur_e = s, e
        else: cur_e = max(cur_e, e)
    if cur_e is not None: tot += cur_e - cur_s
    return tot
busy = union([(s, e) for s, e, _ in ev]); summ = sum(e - s for s, e, _ in ev)
gaps = sorted(((ev[i + 1][0] - max(x[1] for x in ev[:i + 1][-8:])) for i in range(len(ev) - 1)), reverse=True)
print(f"{len(ev)} kernels, span {(t1 - t0) / 1e3:.2f} ms, busy (union) {busy / 1e3:.2f} ms ({100 * busy / (t1 - t0):.1f}%), sum of kernel times {summ / 1e3:.2f} ms "
      f"(overlap {100 * (summ - busy) / summ:.1f}% of kernel time), idle {(t1 - t0 - busy) / 1e3:.2f} ms")
pos = [g for g in gaps if g > 0]
print(f"positive gaps: {len(pos)}, mean {sum(pos) / max(1, len(pos)):.2f} us, total {sum(pos) / 1e3:.2f} ms; largest {[round(g, 1) for g in pos[:8]]}")
