"""ORACLE tooling (test infrastructure only; used only in the build container, where /root/reference exists).

Makes the UNMODIFIED reference package importable (``import pixparse`` from /root/reference/src) although timm,
simple_parsing, chug, jiwer, zss, nltk, Levenshtein and boto3 are not installed and the HF hub is unreachable
(SURVEY.md F5/F6). Every stub carries a real ModuleSpec, and ``transformers`` is imported BEFORE the timm stub
is registered (its ``is_timm_available()`` probe would otherwise raise on ``timm.__spec__ is None``).

What stays real: all of pixparse's own code (Task, Cruller, TextDecoderHf, preprocessing, train_step) and the
installed transformers BartForCausalLM. What is substituted: timm.create_model -> oracle.vit_timm (restated ViT),
timm optim/scheduler/utils -> oracle.timm_helpers (restated), AutoConfig/AutoTokenizer.from_pretrained ->
restated public bart configs / a fake tokenizer with bart's id layout.
"""
import dataclasses
import importlib.machinery
import json
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Serializable:
    """simple_parsing.helpers.Serializable: only .load(json path) -> nested dataclass is used at import time
    (models/config.py:50, tokenizers/config.py:32)."""

    @classmethod
    def load(cls, path):
        with open(path) as f:
            data = json.load(f)
        return _from_dict(cls, data)


def _from_dict(cls, data):
    kwargs = {}
    hints = {f.name: f for f in dataclasses.fields(cls)}
    for k, v in data.items():
        if k not in hints:
            continue
        f = hints[k]
        sub = None
        if f.default_factory is not dataclasses.MISSING and dataclasses.is_dataclass(f.default_factory):
            sub = f.default_factory
        if sub is not None and isinstance(v, dict):
            kwargs[k] = _from_dict(sub, v)
        elif isinstance(v, list):
            kwargs[k] = tuple(v)
        else:
            kwargs[k] = v
    return cls(**kwargs)


class FakeBartTokenizer:
    """Stand-in for AutoTokenizer.from_pretrained('facebook/bart-large') (hub offline): bart's special-token ids
    and vocabulary size, token<->id for added special tokens only (that is all train_step needs)."""

    def __init__(self):
        self.pad_token_id, self.eos_token_id, self.bos_token_id = 1, 2, 0
        self.pad_token, self.eos_token, self.bos_token = "<pad>", "</s>", "<s>"
        self._size = 50265
        self._added = {}

    def __len__(self):
        return self._size

    def add_special_tokens(self, d):
        n = 0
        for t in d.get("additional_special_tokens", []):
            if t not in self._added:
                self._added[t] = self._size
                self._size += 1
                n += 1
        return n

    def add_tokens(self, toks):
        return self.add_special_tokens({"additional_special_tokens": list(toks)})

    def convert_tokens_to_ids(self, t):
        if isinstance(t, (list, tuple)):
            return [self.convert_tokens_to_ids(x) for x in t]
        base = {"<s>": 0, "<pad>": 1, "</s>": 2, "<unk>": 3}
        return self._added.get(t, base.get(t, 3))

    def encode(self, text, add_special_tokens=False):
        return [self.convert_tokens_to_ids(text)]

    def decode(self, ids, **kw):
        inv = {v: k for k, v in self._added.items()}
        return " ".join(inv.get(int(i), str(int(i))) for i in ids)

    def batch_decode(self, batch, **kw):
        return [self.decode(x) for x in batch]


_installed = False


def install():
    """Idempotent. After this, ``import pixparse`` resolves to the reference sources."""
    global _installed
    if _installed:
        return
    import torch
    import transformers  # noqa: F401  (must precede the timm stub)
    from transformers import AutoConfig, AutoTokenizer

    from . import cruller_ref, timm_helpers, vit_timm

    # ---- timm
    timm = _stub("timm", create_model=vit_timm.create_model)
    timm.utils = _stub("timm.utils", NativeScaler=timm_helpers.NativeScaler,
                       dispatch_clip_grad=timm_helpers.dispatch_clip_grad)
    timm.optim = _stub("timm.optim", create_optimizer_v2=timm_helpers.create_optimizer_v2)
    timm.scheduler = _stub("timm.scheduler", create_scheduler_v2=timm_helpers.create_scheduler_v2)
    timm.layers = _stub("timm.layers", SelectAdaptivePool2d=torch.nn.Identity)
    timm.data = _stub("timm.data")
    timm.data.transforms = _stub("timm.data.transforms", CenterCropOrPad=object)
    timm.data.constants = _stub("timm.data.constants", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
                                IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    # ---- simple_parsing
    sp = _stub("simple_parsing", ArgumentParser=object, DashVariant=object, ArgumentGenerationMode=object)
    sp.helpers = _stub("simple_parsing.helpers", Serializable=_Serializable)
    # ---- data / metric libraries that are imported but not exercised by train_step
    chug = _stub("chug", create_wds_loader=None, create_doc_anno_pipe=None, create_image_text_pipe=None)
    chug.common = _stub("chug.common", LoaderBundle=object)
    chug.webdataset = _stub("chug.webdataset", create_doc_anno_pipe=None, create_image_text_pipe=None)
    jiwer = _stub("jiwer", cer=None, wer=None)
    jiwer.transforms = _stub("jiwer.transforms")
    _stub("zss", Node=object)
    _stub("nltk", edit_distance=None)
    _stub("Levenshtein")
    _stub("boto3")
    _stub("albumentations")

    # ---- hub-free config / tokenizer
    def _auto_config(name, *a, **kw):
        return cruller_ref.bart_config(name, dropout_off=_DROPOUT_OFF[0])

    AutoConfig.from_pretrained = staticmethod(_auto_config)
    AutoTokenizer.from_pretrained = staticmethod(lambda name, *a, **kw: FakeBartTokenizer())

    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    _installed = True


_DROPOUT_OFF = [True]


class CpuDeviceEnv:
    """Duck-typed stand-in for pixparse.framework.DeviceEnv, whose __post_init__ asserts CUDA (device.py:112)."""

    def __init__(self, device="cpu"):
        import torch
        self.device = torch.device(device)
        self.world_size, self.local_rank, self.global_rank = 1, 0, 0

    def is_primary(self):
        return True
