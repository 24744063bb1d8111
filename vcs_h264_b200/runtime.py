"""Process-wide default vcs context (one per device), created on first use."""
from __future__ import annotations

from . import _capi

_contexts: dict[int, _capi.Context] = {}


def get_context(device: int = 0) -> _capi.Context:
    ctx = _contexts.get(device)
    if ctx is None or ctx.h is None:
        ctx = _capi.Context(device)      # raises VcsError without a CUDA device: no CPU path
        _contexts[device] = ctx
    return ctx


def close_all():
    for c in _contexts.values():
        c.close()
    _contexts.clear()
