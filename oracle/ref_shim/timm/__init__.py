"""Minimal stand-in for the `timm` package (not installed in this image).

TEST INFRASTRUCTURE ONLY.  The reference package `look2hear.models` imports a
handful of timm symbols at module import time (TDANet.py:9, attentions.py:3,
EMCAD*.py:6-7, TransXNet.py:8-10, swin_*.py); none of them is used by the three
classes on the hot path except `DropPath`, which is the identity in eval mode.
This shim only exists so that `oracle/make_golden.py` can import the unmodified
reference from /root/reference inside the build container.
"""
