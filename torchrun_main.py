"""ReLoRA pre-training entry point — same CLI and YAML surface as the reference ``torchrun_main.py``.

    torchrun --nproc-per-node <N> torchrun_main.py --model_config configs/llama_250m.json \
        --dataset_path <pretokenized dir> --batch_size 24 --total_batch_size 1152 --lr 1e-3 \
        --use_peft --relora 5000 --cycle_length 5000 --restart_warmup_steps 100 \
        --scheduler cosine_restarts --warmup_steps 500 --reset_optimizer_on_relora true \
        --num_training_steps 20000 --save_every 5000 --eval_every 5000 --warmed_up_model <ckpt>

Thin wrapper: parse -> ``relora_b200.engine.run``.  Also runs as a plain ``python torchrun_main.py``
(single process) and on CPU (``--device cpu`` → gloo).
"""
from relora_b200.config import parse_args
from relora_b200.engine import run


def main(argv=None):
    return run(parse_args(argv))


if __name__ == "__main__":
    main()
