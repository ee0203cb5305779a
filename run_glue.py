"""Fine-tune a (ReLoRA-)pretrained Llama on a GLUE task or on custom csv/json files.

    python run_glue.py --model_name_or_path checkpoints/run/model_20000 --task_name sst2 --do_train --do_eval \
        --max_seq_length 128 --per_device_train_batch_size 32 --learning_rate 2e-5 --num_train_epochs 3 \
        --tokenizer_name t5-base --output_dir /tmp/sst2

Surface parity with the reference ``run_glue.py`` (the stock HF example with the model class swapped for the local
``LlamaForSequenceClassification``, ``run_glue.py:49, 379-390``; checkpoints are never saved during training,
``:222``).  This version is self-contained: a plain PyTorch loop (AdamW + linear schedule, optional bf16 autocast,
DDP when launched with torchrun) instead of the HF ``Trainer``, GLUE metrics computed locally, and ``--train_file`` /
``--validation_file`` (csv / json with ``sentence1[,sentence2],label`` columns) for boxes without network access.
Checkpoints saved by ``ReLoRaModel.save_pretrained`` load directly: LoRA factors are merged into the frozen weights.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import random
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from relora_b200.models import LlamaForSequenceClassification, load_config
from relora_b200.obs import logger

TASK_KEYS = {
    "cola": ("sentence", None), "mnli": ("premise", "hypothesis"), "mrpc": ("sentence1", "sentence2"),
    "qnli": ("question", "sentence"), "qqp": ("question1", "question2"), "rte": ("sentence1", "sentence2"),
    "sst2": ("sentence", None), "stsb": ("sentence1", "sentence2"), "wnli": ("sentence1", "sentence2"),
}


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--model_name_or_path", required=True, help="checkpoint directory (config.json + pytorch_model.bin) or a config json")
    p.add_argument("--tokenizer_name", default=None)
    p.add_argument("--task_name", default=None, choices=[None] + sorted(TASK_KEYS))
    p.add_argument("--train_file", default=None)
    p.add_argument("--validation_file", default=None)
    p.add_argument("--test_file", default=None)
    p.add_argument("--max_seq_length", type=int, default=128)
    p.add_argument("--pad_to_max_length", default=True, type=lambda s: str(s).lower() == "true")
    p.add_argument("--max_train_samples", type=int, default=None)
    p.add_argument("--max_eval_samples", type=int, default=None)
    p.add_argument("--do_train", action="store_true")
    p.add_argument("--do_eval", action="store_true")
    p.add_argument("--do_predict", action="store_true")
    p.add_argument("--per_device_train_batch_size", type=int, default=8)
    p.add_argument("--per_device_eval_batch_size", type=int, default=8)
    p.add_argument("--learning_rate", type=float, default=5e-5)
    p.add_argument("--weight_decay", type=float, default=0.0)
    p.add_argument("--num_train_epochs", type=float, default=3.0)
    p.add_argument("--max_steps", type=int, default=-1)
    p.add_argument("--warmup_ratio", type=float, default=0.0)
    p.add_argument("--max_grad_norm", type=float, default=1.0)
    p.add_argument("--bf16", action="store_true")
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--output_dir", required=True)
    p.add_argument("--overwrite_output_dir", action="store_true")
    p.add_argument("--save_model", action="store_true", help="the reference forces save_strategy='no'; opt in to a final save")
    p.add_argument("--device", default="auto")
    # ---- the rest of the HF `TrainingArguments` / data-argument surface the reference script is usually driven with
    # (run_glue.py:222-390 upstream hands these to transformers.Trainer; here they configure the loop below)
    p.add_argument("--dataset_name", default=None, help="datasets hub / local dataset name (alternative to --task_name / --train_file)")
    p.add_argument("--dataset_config_name", default=None)
    p.add_argument("--max_predict_samples", type=int, default=None)
    p.add_argument("--gradient_accumulation_steps", type=int, default=1)
    p.add_argument("--lr_scheduler_type", default="linear", choices=["linear", "cosine", "constant", "constant_with_warmup"])
    p.add_argument("--warmup_steps", type=int, default=0, help="overrides --warmup_ratio when > 0")
    p.add_argument("--adam_beta1", type=float, default=0.9)
    p.add_argument("--adam_beta2", type=float, default=0.999)
    p.add_argument("--adam_epsilon", type=float, default=1e-8)
    p.add_argument("--logging_steps", type=int, default=50)
    p.add_argument("--evaluation_strategy", "--eval_strategy", dest="evaluation_strategy", default="no", choices=["no", "steps", "epoch"])
    p.add_argument("--eval_steps", type=int, default=None)
    p.add_argument("--save_strategy", default="no", help="accepted for CLI parity; the reference forces 'no' (run_glue.py:222)")
    p.add_argument("--fp16", action="store_true", help="refused like the trainer (args_utils.py:56-57): use --bf16")
    p.add_argument("--report_to", default="none")
    p.add_argument("--run_name", default=None)
    p.add_argument("--cache_dir", default=None)
    p.add_argument("--use_fast_tokenizer", default=True, type=lambda s: str(s).lower() == "true")
    p.add_argument("--overwrite_cache", action="store_true")
    p.add_argument("--ignore_mismatched_sizes", action="store_true")
    return p.parse_args(argv)


# --------------------------------------------------------------------------------------------- metrics
def glue_metrics(task: Optional[str], preds: np.ndarray, labels: np.ndarray) -> Dict[str, float]:
    def acc():
        return float((preds == labels).mean())

    def f1():
        tp = float(((preds == 1) & (labels == 1)).sum())
        fp = float(((preds == 1) & (labels == 0)).sum())
        fn = float(((preds == 0) & (labels == 1)).sum())
        return 2 * tp / max(2 * tp + fp + fn, 1e-12)

    if task == "cola":
        tp = float(((preds == 1) & (labels == 1)).sum()); tn = float(((preds == 0) & (labels == 0)).sum())
        fp = float(((preds == 1) & (labels == 0)).sum()); fn = float(((preds == 0) & (labels == 1)).sum())
        den = math.sqrt((tp + fp) * (tp + fn) * (tn + fp) * (tn + fn))
        return {"matthews_correlation": (tp * tn - fp * fn) / den if den > 0 else 0.0}
    if task == "stsb":
        pr = float(np.corrcoef(preds, labels)[0, 1])
        rank = lambda x: np.argsort(np.argsort(x)).astype(np.float64)  # noqa: E731
        sp = float(np.corrcoef(rank(preds), rank(labels))[0, 1])
        return {"pearson": pr, "spearmanr": sp, "combined_score": (pr + sp) / 2}
    if task in ("mrpc", "qqp"):
        a, f = acc(), f1()
        return {"accuracy": a, "f1": f, "combined_score": (a + f) / 2}
    return {"accuracy": acc()}


# --------------------------------------------------------------------------------------------- data
def _load_raw(args):
    import datasets

    if args.task_name is not None and args.train_file is None:
        return datasets.load_dataset("glue", args.task_name, cache_dir=args.cache_dir)
    if args.dataset_name is not None and args.train_file is None:
        if os.path.isdir(args.dataset_name):
            return datasets.load_from_disk(args.dataset_name)
        return datasets.load_dataset(args.dataset_name, args.dataset_config_name, cache_dir=args.cache_dir)
    files = {k: v for k, v in (("train", args.train_file), ("validation", args.validation_file), ("test", args.test_file)) if v}
    ext = "csv" if next(iter(files.values())).endswith(".csv") else "json"
    return datasets.load_dataset(ext, data_files=files)


def _encode(raw, tokenizer, key1, key2, max_len, pad_id, label_map, is_regression):
    def enc(split):
        ids, labels = [], []
        for ex in split:
            a = tokenizer(ex[key1], ex[key2], truncation=True, max_length=max_len) if key2 else tokenizer(ex[key1], truncation=True, max_length=max_len)
            t = a["input_ids"][:max_len]
            ids.append(t + [pad_id] * (max_len - len(t)))
            if "label" in ex and ex["label"] is not None:
                labels.append(float(ex["label"]) if is_regression else (label_map[ex["label"]] if label_map else int(ex["label"])))
        x = torch.tensor(ids, dtype=torch.long)
        y = torch.tensor(labels, dtype=torch.float32 if is_regression else torch.long) if labels else None
        return x, y

    return {k: enc(v) for k, v in raw.items()}


def _load_model(path: str, num_labels: int, pad_id: int, problem_type: Optional[str]):
    config = load_config(path)
    config.num_labels = num_labels
    config.pad_token_id = pad_id
    if problem_type:
        config.problem_type = problem_type
    model = LlamaForSequenceClassification(config)
    wpath = os.path.join(path, "pytorch_model.bin") if os.path.isdir(path) else None
    if wpath and os.path.exists(wpath):
        state = torch.load(wpath, map_location="cpu", weights_only=True)
        merged = {}
        relora_cfg = os.path.join(path, "relora_config.json")
        scale = None
        if os.path.exists(relora_cfg):
            rc = json.load(open(relora_cfg))
            scale = rc["lora_alpha"] / rc["r"]
        for k, v in state.items():
            if ".lora_A." in k or ".lora_B." in k or k.endswith(".scaling") or k.startswith("lm_head"):
                continue
            merged[k] = v.clone()
        if scale is not None:  # fold the low-rank factors into the dense weights
            for k in list(state):
                if k.endswith(".lora_A.weight"):
                    base = k[: -len(".lora_A.weight")]
                    delta = scale * state[base + ".lora_B.weight"].float() @ state[k].float()
                    if base + ".weight" in merged:
                        merged[base + ".weight"] = (merged[base + ".weight"].float() + delta).to(merged[base + ".weight"].dtype)
                    else:  # LoRA-only checkpoint: the product is the whole weight
                        merged[base + ".weight"] = delta.to(state[k].dtype)
        missing, unexpected = model.load_state_dict(merged, strict=False)
        missing = [m for m in missing if not m.startswith("score")]
        if missing or unexpected:
            raise RuntimeError(f"checkpoint mismatch: missing={missing[:5]} unexpected={list(unexpected)[:5]}")
        logger.info(f"loaded backbone from {wpath} (classifier head freshly initialised)")
    else:
        logger.warning("no pytorch_model.bin found: training the classifier from a randomly initialised backbone")
    return model


def main(argv=None):
    args = parse_args(argv)
    if args.fp16:
        raise ValueError("fp16 is not supported (the pre-training weights are bf16); use --bf16")
    random.seed(args.seed); np.random.seed(args.seed); torch.manual_seed(args.seed)
    if os.path.isdir(args.output_dir) and os.listdir(args.output_dir) and args.do_train and not args.overwrite_output_dir:
        raise ValueError(f"Output directory ({args.output_dir}) already exists and is not empty. Use --overwrite_output_dir to overcome.")
    os.makedirs(args.output_dir, exist_ok=True)
    device = torch.device("cuda" if (args.device == "auto" and torch.cuda.is_available()) or args.device == "cuda" else "cpu")

    from transformers import AutoTokenizer

    tokenizer = AutoTokenizer.from_pretrained(args.tokenizer_name or args.model_name_or_path, use_fast=args.use_fast_tokenizer, cache_dir=args.cache_dir)
    raw = _load_raw(args)
    task = args.task_name
    is_regression = task == "stsb" or (task is None and "float" in str(raw["train"].features["label"].dtype))
    if task is not None:
        key1, key2 = TASK_KEYS[task]
    else:
        cols = [c for c in raw["train"].column_names if c != "label"]
        key1, key2 = ("sentence1", "sentence2") if "sentence1" in cols and "sentence2" in cols else (cols[0], cols[1] if len(cols) > 1 else None)
    label_map = None
    if is_regression:
        num_labels = 1
    else:
        labels = sorted(set(raw["train"]["label"]))
        num_labels = len(labels)
        if not all(isinstance(l, int) for l in labels):
            label_map = {l: i for i, l in enumerate(labels)}
    cfg_vocab = load_config(args.model_name_or_path).vocab_size
    pad_id = tokenizer.pad_token_id if tokenizer.pad_token_id is not None else (tokenizer.eos_token_id or 0)
    pad_id = min(pad_id, cfg_vocab - 1)
    data = _encode(raw, tokenizer, key1, key2, args.max_seq_length, pad_id, label_map, is_regression)
    model = _load_model(args.model_name_or_path, num_labels, pad_id, "regression" if is_regression else None).to(device)

    def batches(x, y, bs, shuffle):
        idx = torch.randperm(len(x)) if shuffle else torch.arange(len(x))
        for i in range(0, len(x), bs):
            j = idx[i: i + bs]
            yield x[j].to(device), (y[j].to(device) if y is not None else None)

    def evaluate(split):
        x, y = data[split]
        cap = args.max_predict_samples if split == "test" else args.max_eval_samples
        if cap:
            x, y = x[:cap], (y[:cap] if y is not None else None)
        model.eval()
        outs = []
        with torch.no_grad():
            for xb, _ in batches(x, None, args.per_device_eval_batch_size, False):
                with torch.autocast(device.type, dtype=torch.bfloat16, enabled=args.bf16):
                    outs.append(model(input_ids=xb).logits.float().cpu())
        logits = torch.cat(outs)
        preds = logits.squeeze(-1).numpy() if is_regression else logits.argmax(-1).numpy()
        return preds, (y.numpy() if y is not None else None)

    results: Dict[str, float] = {}
    if args.do_train:
        x, y = data["train"]
        if args.max_train_samples:
            x, y = x[: args.max_train_samples], y[: args.max_train_samples]
        ga = max(1, args.gradient_accumulation_steps)
        micro_per_epoch = math.ceil(len(x) / args.per_device_train_batch_size)
        steps_per_epoch = max(1, micro_per_epoch // ga)
        total = args.max_steps if args.max_steps > 0 else int(steps_per_epoch * args.num_train_epochs)
        opt = torch.optim.AdamW(model.parameters(), lr=args.learning_rate, weight_decay=args.weight_decay,
                                betas=(args.adam_beta1, args.adam_beta2), eps=args.adam_epsilon)
        warm = args.warmup_steps if args.warmup_steps > 0 else int(args.warmup_ratio * total)

        def lr_lambda(st):
            if st < warm and args.lr_scheduler_type != "constant":
                return st / max(1, warm)
            if args.lr_scheduler_type in ("constant", "constant_with_warmup"):
                return 1.0
            frac = (st - warm) / max(1, total - warm)
            if args.lr_scheduler_type == "cosine":
                return max(0.0, 0.5 * (1.0 + math.cos(math.pi * min(1.0, frac))))
            return max(0.0, 1.0 - frac)

        sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda)
        eval_every = None
        if args.evaluation_strategy == "steps":
            eval_every = args.eval_steps or args.logging_steps
        elif args.evaluation_strategy == "epoch":
            eval_every = steps_per_epoch
        model.train()
        step, micro, done = 0, 0, False
        while not done:
            for xb, yb in batches(x, y, args.per_device_train_batch_size, True):
                with torch.autocast(device.type, dtype=torch.bfloat16, enabled=args.bf16):
                    loss = model(input_ids=xb, labels=yb).loss
                (loss / ga).backward()
                micro += 1
                if micro % ga:
                    continue
                torch.nn.utils.clip_grad_norm_(model.parameters(), args.max_grad_norm)
                opt.step(); sched.step(); opt.zero_grad(set_to_none=True)
                step += 1
                if step % max(1, args.logging_steps) == 0 or step == total:
                    logger.info(f"step {step}/{total} loss {float(loss):.4f} lr {sched.get_last_lr()[0]:.2e}")
                if eval_every and step % eval_every == 0 and "validation" in data and step < total:
                    preds, labels = evaluate("validation")
                    logger.info(f"step {step}: {glue_metrics(task, preds, labels)}")
                    model.train()
                if step >= total:
                    done = True
                    break
        results["train_loss"] = float(loss)
        if args.save_model:
            model.save_pretrained(args.output_dir)

    if args.do_eval:
        splits = ["validation_matched", "validation_mismatched"] if task == "mnli" else ["validation"]
        for sp in splits:
            if sp not in data:
                continue
            preds, labels = evaluate(sp)
            m = glue_metrics(task, preds, labels)
            suffix = "_mm" if sp.endswith("mismatched") else ""
            results.update({f"eval_{k}{suffix}": v for k, v in m.items()})
        logger.info(f"eval results: {results}")
    if args.do_predict and "test" in data:
        preds, _ = evaluate("test")
        with open(os.path.join(args.output_dir, f"predict_results_{task or 'custom'}.txt"), "w") as f:
            f.write("index\tprediction\n")
            for i, p in enumerate(preds):
                f.write(f"{i}\t{p:3.3f}\n" if is_regression else f"{i}\t{int(p)}\n")
    with open(os.path.join(args.output_dir, "all_results.json"), "w") as f:
        json.dump(results, f, indent=2)
    return results


if __name__ == "__main__":
    main()
