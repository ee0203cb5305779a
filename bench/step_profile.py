"""Kernel-time breakdown of one training step (torch.profiler, CUDA activities; graphs off so kernels are attributed).

    python bench/step_profile.py [--model llama_250m] [--batch 24] [--seq 512] [--out gpurun_out/step_profile.txt]
"""
import argparse, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RELORA_B200_NO_WANDB", "1")
from relora_b200.engine.api import TrainingEngine
from relora_b200.parallel.dist import DistInfo

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama_250m"); ap.add_argument("--batch", type=int, default=24)
ap.add_argument("--seq", type=int, default=512); ap.add_argument("--ga", type=int, default=2)
ap.add_argument("--out", default=None); ap.add_argument("--engine", default="auto")
a = ap.parse_args()
info = DistInfo(0, 0, 1, torch.device("cuda", 0), "nccl")
torch.cuda.set_device(0)
eng = TrainingEngine.build(info, model_config=os.path.join(ROOT, "configs", f"{a.model}.json"), batch_size=a.batch,
    gradient_accumulation=a.ga, total_batch_size=a.batch * a.ga, max_length=a.seq, use_peft=True, lora_r=128, relora=5000,
    cycle_length=5000, scheduler="cosine_restarts", warmup_steps=500, restart_warmup_steps=100, lr=1e-3,
    num_training_steps=20000, dtype="bfloat16", device="cuda", cuda_graphs=False, engine=a.engine)
ids = torch.randint(0, 32000, (a.ga, a.batch, a.seq), device="cuda")
for _ in range(2):
    eng.train_step_device(ids)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.train_step_device(ids)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = []
for e in ka:
    t = getattr(e, "device_time_total", None)
    if t is None: t = getattr(e, "cuda_time_total", 0)
    if t and e.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((e.key, t, e.count))
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
lines = [f"total device time {tot/1e3:.2f} ms for {a.ga} micro-steps (+1 update) of {a.batch}x{a.seq} tokens, {sum(r[2] for r in rows)} kernels"]
for k, t, c in rows[:45]:
    lines.append(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  x{c:<5d} {k[:150]}")
txt = "\n".join(lines)
print(txt)
if a.out:
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write(txt + "\n")
