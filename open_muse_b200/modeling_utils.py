"""Config / checkpoint surface shared by the models (the reference's "plugin API").

Mirrors the behaviour of ``muse/modeling_utils.py`` that callers rely on -- not its code:
  * ``@register_to_config`` records every ``__init__`` argument (defaults included) in ``self.config``
    and mirrors each one as an attribute on the module *before* ``__init__`` runs
    (reference: muse/modeling_utils.py:827-848,1128-1170) -- this is why ``vq_model.num_embeddings``
    works in training/train_maskgit_imagenet.py:388.
  * ``save_pretrained(dir)`` writes ``config.json`` (sorted keys, indent 2, plus ``_class_name`` and
    ``_version``; :1095-1114) and ``pytorch_model.bin`` (``torch.save(state_dict)``; :331-381).
  * ``from_pretrained(dir, subfolder=, torch_dtype=)`` rebuilds ``cls(**config)`` and loads the state dict,
    returning the model in ``eval()`` (:615).  Checkpoints are interchangeable with the reference's
    because parameter names and shapes are identical.
Hub download is attempted through ``huggingface_hub`` only when the path is not a local directory.
"""
from __future__ import annotations

import functools
import inspect
import json
import os
from typing import Any, Dict, Optional

import torch
from torch import nn

__version__ = "0.0.1"  # value the reference writes into config.json ("_version")

CONFIG_NAME = "config.json"
WEIGHTS_NAME = "pytorch_model.bin"


class Config(dict):
    """dict with attribute access. Attributes assigned after construction (``config.latent_size = ..``)
    are plain attributes, not items, so they are not serialised -- same as the reference's FrozenDict."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in self.items():
            object.__setattr__(self, k, v)

    # the reference's FrozenDict (:770-801) refuses structural edits; item / attribute assignment stays possible there (its
    # frozen flag is never seen through the name mangling), and so it does here
    def _refuse(self, what):
        raise Exception(f"You cannot use ``{what}`` on a {self.__class__.__name__} instance.")

    def __delitem__(self, *args, **kwargs):
        self._refuse("__delitem__")

    def setdefault(self, *args, **kwargs):
        self._refuse("setdefault")

    def pop(self, *args, **kwargs):
        self._refuse("pop")

    def update(self, *args, **kwargs):
        self._refuse("update")


class ConfigMixin:
    config_name = CONFIG_NAME
    ignore_for_config: list = []
    has_compatibles = False

    def register_to_config(self, **kwargs):
        kwargs.pop("kwargs", None)
        for k, v in kwargs.items():
            setattr(self, k, v)
        merged = dict(getattr(self, "_internal_dict", {}))
        merged.update(kwargs)
        self._internal_dict = Config(merged)

    @property
    def config(self) -> Config:
        return self._internal_dict

    # ---- serialisation
    def to_json_string(self) -> str:
        d = {k: v for k, v in self._internal_dict.items()}
        d["_class_name"] = self.__class__.__name__
        d["_version"] = __version__

        def enc(v):
            if hasattr(v, "tolist"):
                return v.tolist()
            if isinstance(v, os.PathLike):
                return str(v)
            return v

        return json.dumps({k: enc(v) for k, v in d.items()}, indent=2, sort_keys=True) + "\n"

    def to_json_file(self, json_file_path):
        """reference :1116-1125"""
        with open(json_file_path, "w", encoding="utf-8") as writer:
            writer.write(self.to_json_string())

    def save_config(self, save_directory, **kwargs):
        if os.path.isfile(save_directory):
            raise AssertionError(f"Provided path ({save_directory}) should be a directory, not a file")
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, self.config_name), "w", encoding="utf-8") as f:
            f.write(self.to_json_string())

    @classmethod
    def load_config(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, **kwargs) -> Dict[str, Any]:
        path = _resolve_file(pretrained_model_name_or_path, cls.config_name, subfolder, **kwargs)
        with open(path, "r", encoding="utf-8") as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config: Dict[str, Any], **kwargs):
        init_kwargs = {k: v for k, v in dict(config).items() if not k.startswith("_")}
        init_kwargs.update(kwargs)
        sig = inspect.signature(cls.__init__)
        if not any(p.kind == inspect.Parameter.VAR_KEYWORD for p in sig.parameters.values()):
            init_kwargs = {k: v for k, v in init_kwargs.items() if k in sig.parameters}
        return cls(**init_kwargs)


def register_to_config(init):
    """Decorator for ``__init__``: positional + keyword + default arguments go to ``register_to_config``."""

    @functools.wraps(init)
    def wrapped(self, *args, **kwargs):
        if not isinstance(self, ConfigMixin):
            raise RuntimeError(f"@register_to_config used on {type(self).__name__}, which is not a ConfigMixin")
        public = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        private = {k: v for k, v in kwargs.items() if k.startswith("_")}
        ignore = set(getattr(self, "ignore_for_config", []))
        params = [(n, p) for i, (n, p) in enumerate(inspect.signature(init).parameters.items()) if i > 0 and n not in ignore]
        record = dict(private)
        for value, (name, _) in zip(args, params):
            record[name] = value
        for name, p in params:
            if name not in record:
                record[name] = public.get(name, p.default)
        self.register_to_config(**record)
        init(self, *args, **public)

    return wrapped


def _resolve_file(name_or_path, filename, subfolder=None, **kwargs) -> str:
    name_or_path = str(name_or_path)
    if os.path.isfile(name_or_path):
        return name_or_path
    if os.path.isdir(name_or_path):
        cand = os.path.join(name_or_path, subfolder, filename) if subfolder else os.path.join(name_or_path, filename)
        if os.path.isfile(cand):
            return cand
        raise EnvironmentError(f"Error no file named {filename} found in directory {name_or_path}.")
    try:  # not a local path: treat as a hub repo id
        from huggingface_hub import hf_hub_download

        return hf_hub_download(name_or_path, filename=filename, subfolder=subfolder,
                               **{k: v for k, v in kwargs.items() if k in ("cache_dir", "revision", "token", "local_files_only")})
    except Exception as e:  # pragma: no cover - no network in tests
        raise EnvironmentError(f"Cannot resolve {filename} for '{name_or_path}': {e}") from e


class ModelMixin(nn.Module):
    config_name = CONFIG_NAME
    _supports_gradient_checkpointing = False

    def __init__(self):
        super().__init__()

    # ---- toggles the training scripts / pipelines call
    @property
    def is_gradient_checkpointing(self) -> bool:
        return bool(getattr(self, "gradient_checkpointing", False))

    def enable_gradient_checkpointing(self):
        if not self._supports_gradient_checkpointing:
            raise ValueError(f"{self.__class__.__name__} does not support gradient checkpointing.")
        self.gradient_checkpointing = True

    def disable_gradient_checkpointing(self):
        if self._supports_gradient_checkpointing:
            self.gradient_checkpointing = False

    def set_use_memory_efficient_attention_xformers(self, valid: bool, attention_op=None) -> None:
        """No-op (reference :276-291 forwards the switch to every child that exposes it)."""

    def enable_xformers_memory_efficient_attention(self, attention_op=None):
        """No-op: attention is always the fused on-chip kernel here (reference :276-329 toggles xformers)."""

    def disable_xformers_memory_efficient_attention(self):
        """No-op, see enable_xformers_memory_efficient_attention."""

    @property
    def device(self) -> torch.device:
        for t in list(self.parameters()) + list(self.buffers()):
            return t.device
        return torch.device("cpu")

    @property
    def dtype(self) -> torch.dtype:
        for t in list(self.parameters()) + list(self.buffers()):
            if t.is_floating_point():
                return t.dtype
        return torch.float32

    def num_parameters(self, only_trainable: bool = False, exclude_embeddings: bool = False) -> int:
        if exclude_embeddings:
            skip = {f"{n}.weight" for n, m in self.named_modules() if isinstance(m, nn.Embedding)}
            params = [p for n, p in self.named_parameters() if n not in skip]
        else:
            params = list(self.parameters())
        return sum(p.numel() for p in params if p.requires_grad or not only_trainable)

    # ---- checkpoints
    def save_pretrained(self, save_directory, is_main_process: bool = True, save_function=None, state_dict=None):
        if os.path.isfile(save_directory):
            raise AssertionError(f"Provided path ({save_directory}) should be a directory, not a file")
        os.makedirs(save_directory, exist_ok=True)
        if is_main_process:
            self.save_config(save_directory)
        if state_dict is None:
            state_dict = self.state_dict()
        if is_main_process:
            (save_function or torch.save)(state_dict, os.path.join(save_directory, WEIGHTS_NAME))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, **kwargs):
        subfolder = kwargs.pop("subfolder", None)
        torch_dtype = kwargs.pop("torch_dtype", None)
        output_loading_info = kwargs.pop("output_loading_info", False)
        kwargs.pop("low_cpu_mem_usage", None)  # both reference modes end in the same state dict load
        kwargs.pop("device_map", None)
        hub_kwargs = {k: kwargs.pop(k) for k in ("cache_dir", "revision", "token", "local_files_only", "force_download",
                                                  "resume_download", "proxies", "use_auth_token") if k in kwargs}
        if torch_dtype is not None and not isinstance(torch_dtype, torch.dtype):
            raise ValueError(f"{torch_dtype} needs to be of type `torch.dtype`, e.g. `torch.float16`.")
        config = cls.load_config(pretrained_model_name_or_path, subfolder=subfolder, **hub_kwargs)
        model = cls.from_config(config, **kwargs)
        weights = _resolve_file(pretrained_model_name_or_path, WEIGHTS_NAME, subfolder, **hub_kwargs)
        state_dict = torch.load(weights, map_location="cpu")
        own = model.state_dict()
        mismatched = [k for k, v in state_dict.items() if k in own and tuple(own[k].shape) != tuple(v.shape)]
        if mismatched:
            raise ValueError(f"Cannot load {weights}: shape mismatch for {mismatched[:5]}")
        result = model.load_state_dict(state_dict, strict=False)
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        model.register_to_config(_name_or_path=str(pretrained_model_name_or_path))
        model.eval()
        if output_loading_info:
            return model, {"missing_keys": list(result.missing_keys), "unexpected_keys": list(result.unexpected_keys),
                           "mismatched_keys": mismatched, "error_msgs": []}
        return model
