"""Headline benchmark: ReLoRA training throughput (tokens/s, whole job) on synthetic tokens.

    python bench.py --gpus 1 --steps 8 --warmup 3                    # this engine
    python bench.py --impl reference --gpus 1 --steps 8 --warmup 3   # unmodified reference (baseline/_ref)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus 8 --steps 8 --warmup 3

Config (BASELINE.json / reference README.md:52-88): llama_250m, ReLoRA r=128, bf16, seq 512, per-GPU
micro-batch 24, gradient accumulation 6 (=> the README's total batch 1152 at 8 GPUs; weak scaling:
global batch 144·N sequences).  One *step* = one optimizer update = 6 micro-batches per GPU, incl.
gradient reduction, clipping, AdamW and the LR schedule.  Random-init weights, synthetic token ids.

Timing: W >= 3 untimed steps, then exactly K steps bracketed by barrier + cuda synchronize, CUDA events on
the launching stream, max over ranks.  The working set of a step (0.5 GB of weights + >1 GB of activations)
exceeds the 126 MB L2, so no explicit flush is needed.  `value` is measured with device-resident inputs;
`e2e.value` goes through the public API (`TrainingEngine.train_step`): pinned-host token ids are copied to
the device every step and the loss is read back to the host every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("RELORA_B200_NO_WANDB", "1")
os.environ.setdefault("WANDB_MODE", "disabled")
os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", type=str, default="llama_250m")
    ap.add_argument("--batch", type=int, default=24, help="per-GPU micro-batch (README.md:55)")
    ap.add_argument("--ga", type=int, default=6, help="gradient accumulation (1152 / (24*8), README.md:56)")
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--lora_r", type=int, default=128)
    ap.add_argument("--engine", type=str, default="auto")
    ap.add_argument("--comm", type=str, default="auto")
    ap.add_argument("--optimizer", type=str, default="adam")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--frozen_dtype", type=str, default=None, help="fp8: E4M3 tensor-core path for the frozen weights (opt-in)")
    ap.add_argument("--attention", type=str, default="auto", choices=["auto", "native", "sdpa"])
    ap.add_argument("--cuda_graphs", type=str, default="true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples `nvidia-smi` clocks / throttle reasons during the timed region (rank 0, GPU 0)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- shared
def setup_dist(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device")
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"[bench] warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    return rank, local, world


def timed(fn_step, steps, world, device):
    """barrier + sync | K steps between CUDA events | sync + barrier; returns max-over-ranks seconds."""
    import torch
    import torch.distributed as dist

    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn_step(i)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def make_tokens(steps, ga, batch, seq, vocab, rank, pinned):
    import torch

    g = torch.Generator().manual_seed(1234 + rank)
    t = torch.randint(0, vocab, (steps, ga, batch, seq), generator=g, dtype=torch.long)
    return t.pin_memory() if pinned else t


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch

    rank, local, world = setup_dist(args)
    from relora_b200.engine.api import TrainingEngine
    from relora_b200.models import load_config
    from relora_b200.ops import native
    from relora_b200.parallel.dist import init_distributed

    info = init_distributed("cuda", "nccl")
    cfg_path = os.path.join(ROOT, "configs", f"{args.model}.json")
    vocab = load_config(cfg_path).vocab_size
    eng = TrainingEngine.build(
        info, model_config=cfg_path, batch_size=args.batch, gradient_accumulation=args.ga,
        total_batch_size=args.batch * args.ga * world, max_length=args.seq, use_peft=True, lora_r=args.lora_r,
        relora=5000, cycle_length=5000, scheduler="cosine_restarts", warmup_steps=500, restart_warmup_steps=100,
        lr=1e-3, num_training_steps=20000, reset_optimizer_on_relora=True, dtype="bfloat16", device="cuda",
        engine=args.engine, comm=args.comm, optimizer=args.optimizer, cuda_graphs=args.cuda_graphs.lower() == "true",
        frozen_dtype=args.frozen_dtype, attention=args.attention,
    )
    dev = info.device
    n_total = args.warmup + args.steps
    host = make_tokens(n_total, args.ga, args.batch, args.seq, vocab, rank, pinned=True)
    dev_tokens = host.to(dev)
    C = native.require()

    for i in range(args.warmup):
        eng.train_step_device(dev_tokens[i])
    if hasattr(eng.stepper, "mark_launch_window"):
        eng.stepper.mark_launch_window()
    else:
        C.reset_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    secs = timed(lambda i: eng.train_step_device(dev_tokens[args.warmup + i]), args.steps, world, dev)
    clocks = sampler.stop() if sampler else None
    launches = int(getattr(eng.stepper, "launches_in_window", lambda n: C.launch_count())(args.steps))
    tokens = eng.tokens_per_step * args.steps
    value = tokens / secs

    e2e = None
    if not args.no_e2e:
        losses = []
        secs2 = timed(lambda i: losses.append(eng.train_step(host[args.warmup + i])), args.steps, world, dev)
        bi = host[0].numel() * host.element_size()
        e2e = {"value": tokens / secs2, "unit": "tokens/s", "h2d_bytes_per_step": bi, "d2h_bytes_per_step": 4,
               "ms_per_step": secs2 / args.steps * 1e3, "last_loss": losses[-1]}
    if rank == 0:
        out = {
            "metric": "training throughput, llama ReLoRA (tokens/s, whole job, device-timed, max over ranks)",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.frozen_dtype not in ("fp8", "fp8_full") else f"bf16 ({args.frozen_dtype}: fp8 tensor-core GEMMs for the frozen weights)",
            "data": "synthetic token ids, random-init weights", "impl": "ours",
            "config": {"model": args.model, "global_batch": args.batch * args.ga * world, "micro_batch_per_gpu": args.batch,
                       "grad_accumulation": args.ga, "seq_len": args.seq, "lora_r": args.lora_r, "lora_dropout": 0.1,
                       "parallelism": f"dp{world}", "optimizer": args.optimizer, "executor": type(eng.stepper).__name__,
                       "comm": getattr(eng.stepper.sync, "transport", "none"),
                       "attention": "tcgen05 (this repo)" if getattr(eng.stepper, "native_attn", False) else "torch SDPA (cuDNN)",
                       "l2": "per-step working set (weights + activations, >1.5 GB) exceeds the 126 MB L2; no flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "native_so": native.so_path(),
        }
        print(json.dumps(out))
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- reference arm
def _stub_missing_modules():
    """The reference imports bitsandbytes at module level (relora.py:10-11); the package is not in this image.
    A stub satisfies the import — the non-quantised code path never touches it.  Nothing in baseline/_ref is edited."""
    for name in ("bitsandbytes", "bitsandbytes.nn", "bitsandbytes.functional"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["bitsandbytes"].nn = sys.modules["bitsandbytes.nn"]
    sys.modules["bitsandbytes"].functional = sys.modules["bitsandbytes.functional"]


def run_reference(args):
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "peft_pretraining")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is not installed (pip --target baseline/_ref /root/reference)"}))
        return
    import torch

    rank, local, world = setup_dist(args)
    import torch.distributed as dist

    _stub_missing_modules()
    sys.path.insert(0, ref_dir)
    try:
        from peft_pretraining import training_utils
        from peft_pretraining.modeling_llama import LlamaForCausalLM
        from peft_pretraining.relora import ReLoRaModel
        from transformers import AutoConfig
    except Exception as e:
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"reference import failed: {type(e).__name__}: {e}"[:300]}))
        dist.destroy_process_group()
        return

    # same steps as the reference's torchrun_main.main (:477-492, 531-553, 598-622, 631-691, 768-826)
    torch.manual_seed(0)
    device = f"cuda:{local}"
    cfg = AutoConfig.from_pretrained(os.path.join(ROOT, "configs", f"{args.model}.json"))
    model = LlamaForCausalLM(cfg)
    model = ReLoRaModel(model, r=args.lora_r, lora_alpha=32, lora_dropout=0.1, target_modules=["attn", "attention", "mlp"],
                        trainable_scaling=False, keep_original_weights=True, lora_only=False, quantize=None, use_double_quant=True)
    model = model.to(device=device, dtype=torch.bfloat16)
    model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], output_device=local)
    trainable = [p for p in model.parameters() if p.requires_grad]
    optimizer = torch.optim.AdamW(trainable, lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999))
    scheduler = training_utils.get_scheculer(optimizer=optimizer, scheduler_type="cosine_restarts", num_training_steps=20000,
                                             warmup_steps=500, min_lr_ratio=0.1, cycle_length=5000, restart_warmup_steps=100,
                                             adjust_step=0)
    ga = args.ga
    n_total = args.warmup + args.steps
    host = make_tokens(n_total, ga, args.batch, args.seq, cfg.vocab_size, rank, pinned=True)
    dev_tokens = host.to(device)
    state = {"loss": 0.0}

    def step(src_tokens):
        loss_info = torch.tensor([0.0, 0.0, 0.0], device=device)
        for mb in range(ga):
            batch = {"input_ids": src_tokens[mb].to(device)}  # no-op for device-resident tokens
            loss = model(**batch, labels=batch["input_ids"]).loss
            loss_info[0] += loss.detach()
            loss_info[1] += 1
            loss_info[2] += torch.isnan(loss).float()
            (loss / ga).backward()
        grad_norm = torch.nn.utils.clip_grad_norm_(trainable, 1.0, error_if_nonfinite=True)
        if rank == 0:
            grad_norm.item()  # the reference logs it to wandb every update (torchrun_main.py:807-808)
        dist.all_reduce(loss_info, op=dist.ReduceOp.SUM)
        _loss = loss_info[0] / loss_info[1]
        if loss_info[2] == 0:
            optimizer.step()
            scheduler.step()
        optimizer.zero_grad()
        return _loss

    for i in range(args.warmup):
        step(dev_tokens[i])
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    secs = timed(lambda i: step(dev_tokens[args.warmup + i]), args.steps, world, torch.device(device))
    clocks = sampler.stop() if sampler else None
    tokens = args.batch * ga * args.seq * world * args.steps
    e2e = None
    if not args.no_e2e:
        def e2e_step(i):
            state["loss"] = float(step(host[args.warmup + i]).item())
        secs2 = timed(e2e_step, args.steps, world, torch.device(device))
        e2e = {"value": tokens / secs2, "unit": "tokens/s", "h2d_bytes_per_step": host[0].numel() * host.element_size(),
               "d2h_bytes_per_step": 4, "ms_per_step": secs2 / args.steps * 1e3, "last_loss": state["loss"]}
    if rank == 0:
        print(json.dumps({
            "metric": "training throughput, llama ReLoRA (tokens/s, whole job, device-timed, max over ranks)",
            "value": tokens / secs, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic token ids, random-init weights", "impl": "reference",
            "config": {"model": args.model, "global_batch": args.batch * ga * world, "micro_batch_per_gpu": args.batch,
                       "grad_accumulation": ga, "seq_len": args.seq, "lora_r": args.lora_r, "lora_dropout": 0.1,
                       "parallelism": f"dp{world}", "optimizer": "adam", "executor": "reference torch eager + DDP"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": 0,
        }))
    dist.barrier()
    dist.destroy_process_group()


class _QuietStdout:
    """Route everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr so that stdout carries exactly
    the one JSON line of the bench contract."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        real = os.fdopen(os.dup(self._saved), "w")
        self._print = builtins_print = print

        def emit(*a, **k):
            builtins_print(*a, **dict(k, file=real, flush=True))

        globals()["print"] = emit
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        globals().pop("print", None)
        return False


if __name__ == "__main__":
    a = parse()
    with _QuietStdout():
        if a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)
