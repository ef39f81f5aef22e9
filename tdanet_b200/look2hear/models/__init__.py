"""look2hear.models mirror: the TDANet classes on the hot path (TDANetBest, the fork TDANet, TDANetMultRes, and the
TDANetOrigin / TDANetYang heads of configs/tdanet_origin.yml / tdanet.yml), `get`, `register_model`
(reference: look2hear/models/__init__.py:82-114)."""
from .base_model import BaseModel
from .tdanet import TDANet, TDANetBest, TDANetMultRes, TDANetOrigin, TDANetYang

__all__ = ["BaseModel", "TDANet", "TDANetBest", "TDANetMultRes", "TDANetOrigin", "TDANetYang"]


def register_model(custom_model):
    """Register a custom model class, gettable with `models.get`."""
    name = custom_model.__name__
    if name in globals() or name.lower() in globals():
        raise ValueError(f"Model {name} already exists. Choose another name.")
    globals()[name] = custom_model


def get(identifier):
    """Model class from its (case-insensitive) name."""
    if isinstance(identifier, str):
        found = {k.lower(): v for k, v in globals().items()}.get(identifier.lower())
        if found is not None:
            return found
    raise ValueError(f"Could not interpret model name : {str(identifier)}")
