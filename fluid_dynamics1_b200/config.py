"""The reference's config.txt system (src/config.c) through the C ABI."""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import Config


def load_default_config() -> Config:
    c = Config()
    _lib.lib().cnv_config_default(C.byref(c))
    return c


def load_config_from_file(filename: str) -> Config:
    c = Config()
    _lib.lib().cnv_config_from_file(filename.encode(), C.byref(c))
    return c


def print_config(cfg: Config) -> None:
    _lib.lib().cnv_config_print(C.byref(cfg))


def config_from_dict(d: dict) -> Config:
    c = load_default_config()
    for k, v in d.items():
        if not hasattr(c, k):
            raise KeyError(f"Warning: Unknown configuration parameter '{k}'")
        setattr(c, k, v)
    return c
