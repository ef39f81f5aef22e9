"""BaseModel mirror (reference: look2hear/models/base_model.py:111-181).

Same public surface: sample_rate(), from_pretrain(model_name, path_or_hub_id, *args, **kwargs),
serialize(), get_state_dict(), get_model_args(), load_state_dict_in_audio().
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

CACHE_DIR = os.path.expanduser("~/.cache/torch/tdanet")


def _unsqueeze_to_3d(x):
    if x.ndim == 1:
        return x.reshape(1, 1, -1)
    if x.ndim == 2:
        return x.unsqueeze(1)
    return x


class BaseModel(nn.Module):
    def __init__(self, sample_rate, in_chan=1):
        super().__init__()
        self._sample_rate = sample_rate
        self._in_chan = in_chan

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def sample_rate(self):
        return self._sample_rate

    @staticmethod
    def load_state_dict_in_audio(model, pretrained_dict):
        merged = model.state_dict()
        for key, value in pretrained_dict.items():
            if "audio_model" in key:
                merged[key[len("audio_model."):]] = value
        model.load_state_dict(merged)
        return model

    @staticmethod
    def from_pretrain(model_name, pretrained_model_conf_or_path, *args, **kwargs):
        """Local file: build `model_name(*args, **kwargs)`, strip the Lightning `audio_model.` prefix and
        load non-strictly.  Otherwise a hub id / URL holding {model_name, model_args, state_dict}."""
        from . import get

        if os.path.exists(pretrained_model_conf_or_path):
            conf = torch.load(pretrained_model_conf_or_path, map_location="cpu")
            model = get(model_name)(*args, **kwargs)
            state = OrderedDict((k.replace("audio_model.", ""), v) for k, v in conf["state_dict"].items())
            model.load_state_dict(state, strict=False)
            return model
        path = _cached_download(pretrained_model_conf_or_path)
        conf = torch.load(path, map_location="cpu")
        model = get(conf["model_name"])(*args, **conf["model_args"])
        model.load_state_dict(conf["state_dict"])
        return model

    def serialize(self):
        infos = {"software_versions": {"torch_version": torch.__version__}}
        try:  # the reference records the Lightning version; optional here
            import pytorch_lightning as pl
            infos["software_versions"]["pytorch_lightning_version"] = pl.__version__
        except ImportError:
            pass
        return dict(model_name=self.__class__.__name__, state_dict=self.get_state_dict(),
                    model_args=self.get_model_args(), infos=infos)

    def get_state_dict(self):
        return self.state_dict()

    def get_model_args(self):
        raise NotImplementedError


def _cached_download(name_or_url):
    """Hub ids / URLs need the network; resolved lazily so that offline use never imports hub code."""
    import huggingface_hub

    if name_or_url.startswith(("http://", "https://")):
        from hashlib import sha256
        from torch import hub

        target_dir = os.path.join(CACHE_DIR, sha256(name_or_url.encode("utf-8")).hexdigest())
        os.makedirs(target_dir, exist_ok=True)
        target = os.path.join(target_dir, "model.pth")
        if not os.path.isfile(target):
            hub.download_url_to_file(name_or_url, target)
        return target
    model_id, _, revision = name_or_url.partition("@")
    os.makedirs(CACHE_DIR, exist_ok=True)
    return huggingface_hub.hf_hub_download(repo_id=model_id, filename="pytorch_model.bin",
                                           cache_dir=CACHE_DIR, revision=revision or None)
