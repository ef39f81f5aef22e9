"""Mirror of the reference `look2hear` package surface for the separation hot path.

`tdanet_b200.look2hear.models` / `.losses` / `.system` keep the reference's class names, constructor
kwargs, forward signatures and state_dict keys, so `import tdanet_b200.look2hear as
look2hear` is the whole migration for code that only touches this path.
"""
from . import losses, metrics, models, system  # noqa: F401
from .._lib import deterministic, set_deterministic  # noqa: F401  (bit-reproducible inference: exact GlobLN sums)
