"""NeoX/Megatron data path: native index builders (golden values from SURVEY.md Appendix B), file formats, loaders."""
import os
import struct

import numpy as np
import pytest
import torch
import yaml

from relora_b200.data.neox import (BlendableDataset, DistributedBatchSampler, MMapIndexedDataset, MMapIndexedDatasetBuilder,
                                   NeoXArgs, build_train_valid_test_dataloaders, make_builder, make_dataset)
from relora_b200.data.neox import gpt2_dataset as g2
from relora_b200.data.neox import helpers_build
from relora_b200.data.neox.data_utils import (get_normalized_weights_and_num_samples, get_train_valid_test_split_,
                                              weights_by_num_docs)
from relora_b200.data.neox.indexed_dataset import IndexedDataset, IndexedDatasetBuilder, infer_dataset_impl


@pytest.fixture(scope="module")
def helpers():
    return helpers_build.load()


SIZES = np.array([5, 3, 7, 2, 9, 4], dtype=np.int32)
DOC_IDX = np.array([2, 0, 5, 1, 4, 3, 3, 1, 0, 4, 2, 5], dtype=np.int32)
GOLD_SAMPLE = [[0, 0], [0, 4], [1, 1], [2, 0], [3, 0], [4, 1], [4, 5], [5, 0], [7, 0], [8, 1], [9, 0], [9, 4], [9, 8], [10, 3], [11, 0]]


def test_build_sample_idx_golden(helpers):
    a = helpers.build_sample_idx_int32(SIZES, DOC_IDX, 4, 2, 30)
    b = helpers.build_sample_idx_int64(SIZES, DOC_IDX, 4, 2, 30)
    assert a.dtype == np.int32 and b.dtype == np.int64 and a.shape == (15, 2)
    assert a.tolist() == GOLD_SAMPLE and b.tolist() == GOLD_SAMPLE
    assert g2.build_sample_idx_python(SIZES, DOC_IDX, 4, 2, 30).tolist() == GOLD_SAMPLE


def test_build_sample_idx_random_matches_python(helpers):
    rng = np.random.RandomState(0)
    sizes = rng.randint(1, 50, size=200).astype(np.int32)
    docs = np.arange(200, dtype=np.int32)
    doc_idx = g2.build_doc_idx(docs, 3, rng)
    tpe = int(sizes.sum())
    a = helpers.build_sample_idx_int32(sizes, doc_idx, 17, 3, tpe)
    assert a.tolist() == g2.build_sample_idx_python(sizes, doc_idx, 17, 3, tpe).tolist()


def test_build_blending_indices_golden(helpers):
    di, ds = np.zeros(20, dtype=np.uint8), np.zeros(20, dtype=np.int64)
    helpers.build_blending_indices(di, ds, np.array([0.5, 0.3, 0.2]), 3, 20, False)
    assert di.tolist() == [0, 1, 2, 0, 1, 0, 2, 0, 1, 0, 0, 1, 2, 0, 1, 0, 2, 0, 1, 0]
    assert ds.tolist() == [0, 0, 0, 1, 1, 2, 1, 3, 2, 4, 5, 3, 2, 6, 4, 7, 3, 8, 5, 9]


def test_build_mapping_and_blocks_golden(helpers):
    docs = np.array([0, 3, 5, 9], dtype=np.int64)
    sizes = np.array([10, 20, 30, 5, 5, 100, 100, 100, 100], dtype=np.int32)
    m = helpers.build_mapping(docs, sizes, 2, 1000, 128, 0.1, 1234, False)
    assert m.dtype == np.uint32
    assert m.tolist() == [[0, 3, 128], [3, 5, 128], [3, 5, 128], [7, 9, 128], [0, 3, 128], [5, 7, 128], [5, 7, 128], [7, 9, 128]]
    b = helpers.build_blocks_mapping(docs, sizes, np.array([2, 2, 2], dtype=np.int32), 1, 1000, 64, 1234, False, False)
    assert b.tolist() == [[5, 7, 2, 2], [0, 3, 0, 0], [3, 5, 1, 1], [7, 9, 2, 3]]


def _write_corpus(prefix, n_docs=40, vocab=1000, seed=0, impl="mmap"):
    rng = np.random.RandomState(seed)
    b = make_builder(prefix + ".bin", impl, vocab_size=vocab)
    docs = []
    for _ in range(n_docs):
        d = rng.randint(0, vocab, size=rng.randint(5, 60))
        docs.append(d)
        b.add_item(d)
        b.end_document()
    b.finalize(prefix + ".idx")
    return docs


def test_mmap_format_and_reader(tmp_path):
    prefix = str(tmp_path / "corpus")
    docs = _write_corpus(prefix)
    raw = open(prefix + ".idx", "rb").read()
    assert raw[:9] == b"MMIDIDX\x00\x00"
    assert struct.unpack("<Q", raw[9:17]) == (1,) and raw[17] == 8  # version 1, dtype code 8 = uint16
    n, nd = struct.unpack("<QQ", raw[18:34])
    assert n == 40 and nd == 41
    assert infer_dataset_impl(prefix) == "mmap"
    ds = make_dataset(prefix, "mmap", skip_warmup=True)
    assert isinstance(ds, MMapIndexedDataset) and len(ds) == 40
    assert ds.sizes.tolist() == [len(d) for d in docs]
    for i in (0, 7, 39):
        assert ds[i].tolist() == docs[i].tolist()
    assert ds.get(3, offset=2, length=3).tolist() == docs[3][2:5].tolist()
    assert [x.tolist() for x in ds[2:5]] == [d.tolist() for d in docs[2:5]]
    assert ds.doc_idx.tolist() == list(range(41))


def test_legacy_format_roundtrip(tmp_path):
    prefix = str(tmp_path / "legacy")
    b = IndexedDatasetBuilder(prefix + ".bin", dtype=np.int32)
    docs = [np.arange(5), np.arange(3) + 10, np.arange(7) + 20]
    for d in docs:
        b.add_item(d)
        b.end_document()
    b.finalize(prefix + ".idx")
    assert open(prefix + ".idx", "rb").read(8) == b"TNTIDX\x00\x00"
    ds = IndexedDataset(prefix)
    assert len(ds) == 3 and ds[1].tolist() == docs[1].tolist()
    assert infer_dataset_impl(prefix) == "cached"
    cached = make_dataset(prefix, "cached")
    cached.prefetch([0, 2])
    assert cached[2].tolist() == docs[2].tolist()


def test_gpt2_dataset_samples_are_contiguous_text(tmp_path):
    prefix = str(tmp_path / "corpus")
    docs = _write_corpus(prefix, n_docs=30, seed=3)
    indexed = make_dataset(prefix, "mmap", skip_warmup=True)
    ds = g2.GPT2Dataset("train", prefix, np.arange(30, dtype=np.int32), indexed, num_samples=50, seq_length=16, seed=1234)
    assert len(ds) >= 50
    for f in ("doc", "sample", "shuffle"):
        assert os.path.exists(f"{prefix}_train_indexmap_50ns_16sl_1234s_{f}_idx.npy")
    # every sample has seq_length + 1 tokens and equals the corresponding slice of the shuffled document stream
    stream = np.concatenate([docs[d] for d in ds.doc_idx])
    for i in range(len(ds)):
        item = ds[i]["input_ids"]
        assert item.dtype == np.int64 and item.shape == (17,)
        s = int(ds.shuffle_idx[i])
        assert item.tolist() == stream[s * 16: s * 16 + 17].tolist()
    # rebuilding loads the cached maps (same object content)
    ds2 = g2.GPT2Dataset("train", prefix, np.arange(30, dtype=np.int32), indexed, num_samples=50, seq_length=16, seed=1234)
    assert np.array_equal(ds2.shuffle_idx, ds.shuffle_idx)
    assert ds[len(ds) + 3]["input_ids"].tolist() == ds[3]["input_ids"].tolist()  # modulo wrap on overflow


def test_blendable_dataset_and_weights(tmp_path):
    class Const(torch.utils.data.Dataset):
        def __init__(self, v, n):
            self.v, self.n = v, n

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            if i >= self.n:
                raise IndexError
            return (self.v, i)

    b = BlendableDataset([Const(0, 50), Const(1, 30), Const(2, 20)], [5, 3, 2])
    assert len(b) == 100
    got = [b[i][0] for i in range(100)]
    assert abs(got.count(0) - 50) <= 1 and abs(got.count(1) - 30) <= 1
    w, n = get_normalized_weights_and_num_samples([1.0, 3.0], 1000)
    assert w == [0.25, 0.75] and n == [252, 754]
    assert get_train_valid_test_split_("969, 30, 1", 1000) == [0, 969, 999, 1000]
    assert get_train_valid_test_split_("8/1/1", 10) == [0, 8, 9, 10]
    assert weights_by_num_docs([100]) == [1.0]
    ws = weights_by_num_docs([1000, 10], alpha=0.3)
    assert abs(sum(ws) - 1) < 1e-9 and ws[1] > 10 / 1010


def test_distributed_batch_sampler_resume():
    data = list(range(40))
    s = torch.utils.data.SequentialSampler(data)
    a = DistributedBatchSampler(s, batch_size=8, drop_last=True, rank=1, world_size=2)
    assert list(a) == [[4, 5, 6, 7], [12, 13, 14, 15], [20, 21, 22, 23], [28, 29, 30, 31], [36, 37, 38, 39]]
    a.start_iter = 3
    assert list(a) == [[28, 29, 30, 31], [36, 37, 38, 39]]
    assert a.start_iter == 0
    inter = DistributedBatchSampler(s, batch_size=8, drop_last=True, rank=0, world_size=2, interleave=True)
    assert next(iter(inter)) == [0, 2, 4, 6]


def test_neox_args_and_dataloaders_end_to_end(tmp_path):
    prefix = str(tmp_path / "pile")
    _write_corpus(prefix, n_docs=200, seed=5)
    conf = {"data-path": prefix, "split": "8,1,1", "data-impl": "mmap", "seq-length": 32, "train-iters": 20, "eval-interval": 10,
            "eval-iters": 2, "global_num_gpus": 1, "train_micro_batch_size_per_gpu": 4, "gradient_accumulation_steps": 2,
            "num-workers": 0, "hidden-size": 64, "deepspeed": True}
    args = NeoXArgs.from_dict(conf)
    assert args.train_batch_size == 8 and args.batch_size == 4 and args.seed == 1234
    assert NeoXArgs.calculate_batch_parameters(4, train_batch=64, micro_batch=4) == (64, 4, 4)
    assert NeoXArgs.calculate_batch_parameters(2, micro_batch=3, grad_acc=5) == (30, 3, 5)
    with pytest.raises(RuntimeError):
        NeoXArgs.from_dict({k: v for k, v in conf.items() if k != "global_num_gpus"})
    tl, vl, sl = build_train_valid_test_dataloaders(args)
    batch = next(iter(tl))
    assert batch["input_ids"].shape == (4, 33) and batch["input_ids"].dtype == torch.int64  # seq_length + 1 tokens
    assert args.do_train == 1 and vl is not None and sl is not None
    tl.batch_sampler.start_iter = 5
    resumed = next(iter(tl))
    it = iter(build_train_valid_test_dataloaders(NeoXArgs.from_dict(conf))[0])
    for _ in range(5):
        next(it)
    assert torch.equal(resumed["input_ids"], next(it)["input_ids"])


def test_trainer_with_megatron_dataset_config(tmp_path):
    from torchrun_main import main

    prefix = str(tmp_path / "pile")
    _write_corpus(prefix, n_docs=400, vocab=32000, seed=9)
    y = tmp_path / "data.yaml"
    y.write_text(yaml.safe_dump({"data-path": prefix, "split": "8,1,1", "data-impl": "mmap", "seq-length": 31, "train-iters": 10,
                                 "eval-interval": 5, "eval-iters": 1}))
    cfg = os.path.join(os.path.dirname(os.path.dirname(__file__)), "configs", "llama_9m.json")
    res = main(["--model_config", cfg, "--megatron_dataset_config", str(y), "--batch_size", "2", "--total_batch_size", "4",
                "--max_length", "31", "--lr", "1e-3", "--scheduler", "cosine", "--warmup_steps", "1", "--num_training_steps", "4",
                "--save_every", "100", "--eval_every", "100", "--save_dir", str(tmp_path / "run"), "--device", "cpu",
                "--dtype", "float32", "--workers", "0"])
    assert res["update_step"] == 4 and "final_test_loss" in res


def _tiny_pythia_dir(path, vocab=512, hidden=64, heads=4):
    """A local Pythia-style checkpoint directory: HF ``config.json`` + ``pytorch_model.bin`` with HF GPT-NeoX key names."""
    transformers = pytest.importorskip("transformers")
    import json

    kw = dict(vocab_size=vocab, hidden_size=hidden, num_hidden_layers=2, num_attention_heads=heads, intermediate_size=4 * hidden,
              rotary_pct=0.25, max_position_embeddings=64, use_parallel_residual=True, hidden_act="gelu", layer_norm_eps=1e-5)
    torch.manual_seed(0)
    hf = transformers.GPTNeoXForCausalLM(transformers.GPTNeoXConfig(**kw))
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(dict(model_type="gpt_neox", **kw), f)
    torch.save(hf.state_dict(), os.path.join(path, "pytorch_model.bin"))
    return path


def _run_pythia_recipe(tmp_path, device, dtype, hidden=64):
    """The reference's only shipped recipe (training_configs/1B_v1.0.yaml: Pythia warm start + Megatron data + ReLoRA with magnitude
    pruning, ``force_keep_original``, AdamW betas / weight decay) at toy scale, through the YAML front end, then auto-resumed."""
    from torchrun_main import main

    ckpt = _tiny_pythia_dir(str(tmp_path / "pythia-tiny"), hidden=hidden)
    prefix = str(tmp_path / "pile")
    _write_corpus(prefix, n_docs=400, vocab=512, seed=11)
    data = tmp_path / "data.yaml"
    data.write_text(yaml.safe_dump({"data-path": prefix, "split": "8,1,1", "data-impl": "mmap", "seq-length": 32, "train-iters": 20,
                                    "eval-interval": 5, "eval-iters": 1}))
    recipe = tmp_path / "recipe.yaml"
    recipe.write_text(yaml.safe_dump(dict(
        model_name_or_path=ckpt, model_revision="step1000", dtype=dtype, distributed_type="ddp",
        megatron_dataset_config=str(data), max_length=32, workers=0,
        use_peft=True, lora_r=8, relora=3, force_keep_original=True, restart_warmup_steps=1, reset_optimizer_on_relora=False,
        optimizer_magnitude_pruning=0.8, optimizer="adam", lr=4e-4, adam_beta1=0.9, adam_beta2=0.95, weight_decay=0.01,
        scheduler="cosine_restarts", warmup_steps=2, batch_size=2, total_batch_size=4, num_training_steps=9,
        save_dir=str(tmp_path / "run"), autoresume=True, save_every=3, eval_every=3, tags="relora1b", comment="toy")))
    res = main(["--training_config", str(recipe), "--device", device])
    assert res["update_step"] == 9 and res["n_lora_restarts"] == 2 and res["n_optimizer_resets"] == 2
    assert "final_test_loss" in res and res["final_eval_loss"] == res["final_eval_loss"]  # finite
    # the batches carry seq_length + 1 tokens (Megatron convention): the model ran at T = 33
    assert res["tokens_seen"] == 9 * 4 * 33
    # extend the run: autoresume picks up the last checkpoint and continues with the same data order
    recipe.write_text(recipe.read_text().replace("num_training_steps: 9", "num_training_steps: 12"))
    res2 = main(["--training_config", str(recipe), "--device", device])
    assert res2["update_step"] == 12 and res2["n_lora_restarts"] == 3
    return res, res2


def test_pythia_recipe_end_to_end(tmp_path):
    _run_pythia_recipe(tmp_path, "cpu", "float32")


@pytest.mark.gpu
def test_pythia_recipe_end_to_end_gpu(tmp_path):
    """Same recipe on the device: bf16, this repo's LayerNorm / GELU / partial-rotary / LoRA-linear kernels, odd sequence length (33)."""
    res, res2 = _run_pythia_recipe(tmp_path, "cuda", "bfloat16", hidden=128)
    assert res["final_eval_loss"] < 7.5 and res2["final_eval_loss"] < 7.5   # ln(512) = 6.24 at init
