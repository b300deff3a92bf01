"""Harness that runs the UNMODIFIED reference `FullModel`, training epoch and caption generation with this
repository's `BridgeLite` swapped in (BASELINE.json configs[1] "full training step ... frozen
DINOv2-large + Gemma-2-2B random-init") -- TEST / BENCH INFRASTRUCTURE.

The reference package is imported from oracle/_ref (copied there byte for byte by oracle/make_ref.py; present in the
build container and on the GPU box, never committed). There is no network and there are no checkpoints, so the
frozen models are random-initialised from their configs: `Gemma2Config()` defaults are exactly gemma-2-2b (hidden
2304, 26 layers, vocabulary 256000) and `Dinov2Config(hidden_size=1024, num_hidden_layers=24,
num_attention_heads=16, image_size=518, patch_size=14)` is DINOv2-large. `from_pretrained` is patched IN THE
NAMESPACES the reference binds at import time (vision_encoder.py:16, language_model.py:17) to return these models
and stub tokenizer / image processor objects; the swap-in point is the name `BridgeLite` in
`vlm_bridge.model_architecture.full_model` (imported at full_model.py:22), exactly what INTEGRATION.md tells a
user to patch. Nothing of the reference is modified.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.path.join(ROOT, "oracle", "_ref")


def available() -> tuple[bool, str]:
    if not os.path.isdir(os.path.join(REF_ROOT, "vlm_bridge")):
        return False, "oracle/_ref/vlm_bridge absent (run oracle/make_ref.py in the build container)"
    try:
        import transformers  # noqa: F401
    except Exception as e:  # noqa: BLE001
        return False, f"transformers not importable: {e!r}"
    return True, ""


class StubTokenizer:
    """Gemma special ids (data_loader.py:325: BOS 2, PAD 0; EOS 1); decoding = the ids as decimal text, so that the
    ids can be recovered from the caption string `generate_caption` returns."""
    bos_token, eos_token, pad_token = "<bos>", "<eos>", "<pad>"
    bos_token_id, eos_token_id, pad_token_id = 2, 1, 0

    def batch_decode(self, token_ids, skip_special_tokens=True):
        return [" ".join(str(int(t)) for t in row) for row in token_ids]

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(str(int(t)) for t in ids)


class StubImageProcessor:
    def __call__(self, images, return_tensors="pt"):
        return {"pixel_values": torch.stack([torch.as_tensor(i) for i in images])}


def import_reference():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    sys.dont_write_bytecode = True
    import vlm_bridge.model_architecture.full_model as fm
    import vlm_bridge.model_architecture.language_model as lm
    import vlm_bridge.model_architecture.vision_encoder as ve
    return fm, lm, ve


def build_full_model(bridge_cls=None, device: str = "cuda", gemma_layers: int | None = None,
                     dino_layers: int | None = None, bridge_dropout: float = 0.1, seed: int = 0):
    """The reference `FullModel(device=...)` over random-init frozen models. `bridge_cls` replaces the name
    `BridgeLite` in the reference's full_model namespace for the duration of the construction (None = the
    reference's own). `gemma_layers` / `dino_layers` reduce the depth of the frozen models (tests; the bridge and
    every width stay at their real sizes); None = the real 26 / 24 layers."""
    from transformers import Dinov2Config, Dinov2Model, Gemma2Config, Gemma2ForCausalLM

    fm, lm, ve = import_reference()
    gcfg = Gemma2Config()
    if gemma_layers is not None:
        gcfg.num_hidden_layers = gemma_layers
        if hasattr(gcfg, "layer_types") and gcfg.layer_types is not None:
            gcfg.layer_types = list(gcfg.layer_types)[:gemma_layers]
    dcfg = Dinov2Config(hidden_size=1024, num_hidden_layers=24 if dino_layers is None else dino_layers,
                        num_attention_heads=16, image_size=518, patch_size=14)

    def make_dino(name, **kw):
        torch.manual_seed(seed + 1)
        with torch.device(device):
            return Dinov2Model(dcfg)

    def make_gemma(name, device_map=None, **kw):
        torch.manual_seed(seed + 2)
        with torch.device(device):
            return Gemma2ForCausalLM(gcfg)

    saved = (ve.AutoModel, ve.AutoImageProcessor, lm.AutoModelForCausalLM, lm.AutoTokenizer, fm.BridgeLite)
    ve.AutoModel = types.SimpleNamespace(from_pretrained=make_dino)
    ve.AutoImageProcessor = types.SimpleNamespace(from_pretrained=lambda name, **kw: StubImageProcessor())
    lm.AutoModelForCausalLM = types.SimpleNamespace(from_pretrained=make_gemma)
    lm.AutoTokenizer = types.SimpleNamespace(from_pretrained=lambda name, **kw: StubTokenizer())
    if bridge_cls is not None:
        fm.BridgeLite = bridge_cls
    try:
        torch.manual_seed(seed)
        model = fm.FullModel(bridge_dropout=bridge_dropout, device=device)
    finally:
        ve.AutoModel, ve.AutoImageProcessor, lm.AutoModelForCausalLM, lm.AutoTokenizer, fm.BridgeLite = saved
    return model


def reference_bridge_cls():
    fm, _, _ = import_reference()
    import vlm_bridge.model_architecture.bridge_module as bm
    return bm.BridgeLite


def make_batches(n: int, batch: int, length: int, image: int = 224, seed: int = 1234, vocab: int = 256000):
    """images randn [B, 3, image, image]; ids randint(3, vocab) with BOS first; mask of ones."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        ids = torch.randint(3, vocab, (batch, length), generator=g)
        ids[:, 0] = 2
        out.append({"images": torch.randn(batch, 3, image, image, generator=g), "input_ids": ids,
                    "attention_mask": torch.ones(batch, length, dtype=torch.long)})
    return out


class _Writer:
    def __init__(self):
        self.scalars = []

    def add_scalar(self, tag, value, step):
        self.scalars.append((tag, float(value), int(step)))


def training_context(model, batches, lr: float = 1e-5, clip: float = 0.3, optimizer=None):
    """A `TrainingContext` for the reference's run_training_epoch (core_training_loop.py:16-134): AdamW as
    training_setup.py:248-254 builds it, GradScaler + bf16 autocast as configure_hardware_and_precision does on CUDA
    (training_setup.py:215-218), clip 0.3 (config/training-default.yaml:9)."""
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from vlm_bridge.training_strategy.training_setup import TrainingConfig, TrainingContext

    cfg = TrainingConfig()
    cfg.use_amp, cfg.amp_dtype, cfg.gradient_clip_val, cfg.log_every_n_steps = True, "bfloat16", clip, 1
    if optimizer is None:
        optimizer = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=lr, weight_decay=0.01,
                                      betas=(0.9, 0.999), eps=1e-8)
    dev = torch.device(model.device)
    return TrainingContext(config=cfg, model=model, optimizer=optimizer, scheduler=None, train_loader=batches,
                           val_loader=[], device=dev, scaler=torch.amp.GradScaler("cuda"), writer=_Writer(),
                           checkpoint_dir=None)


@contextlib.contextmanager
def quiet():
    with open(os.devnull, "w") as f, contextlib.redirect_stdout(f), contextlib.redirect_stderr(f):
        yield


def ids_from_caption(caption: str) -> list[int]:
    return [int(t) for t in caption.split()]
