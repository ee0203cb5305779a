from .configs import SimpleConfig, config_to_dict, load_config, save_config
from .llama import (
    CausalLMOutput,
    LlamaDecoderLayer,
    LlamaForCausalLM,
    LlamaForSequenceClassification,
    LlamaModel,
    LlamaRMSNorm,
)
from .pythia import GPTNeoXForCausalLM, GPTNeoXLayer, GPTNeoXModel


def build_causal_lm(config):
    """Instantiate the causal LM class matching ``config.model_type`` (random init)."""
    mt = getattr(config, "model_type", "llama")
    if mt == "llama":
        return LlamaForCausalLM(config)
    if mt == "gpt_neox":
        return GPTNeoXForCausalLM(config)
    raise NotImplementedError(f"Unknown model config type {mt}, only LLaMA and GPT-NeoX are supported")


def model_from_config_dir(path: str):
    """``AutoModelForCausalLM.from_config(AutoConfig.from_pretrained(path))`` equivalent."""
    return build_causal_lm(load_config(path))


__all__ = [
    "SimpleConfig",
    "load_config",
    "save_config",
    "config_to_dict",
    "CausalLMOutput",
    "LlamaForCausalLM",
    "LlamaForSequenceClassification",
    "LlamaModel",
    "LlamaDecoderLayer",
    "LlamaRMSNorm",
    "GPTNeoXForCausalLM",
    "GPTNeoXModel",
    "GPTNeoXLayer",
    "build_causal_lm",
    "model_from_config_dir",
]
