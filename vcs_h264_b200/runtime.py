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


# Exclusive contexts for the clip front-ends: a ClipEncoder/ClipDecoder owns its context (Q tables, stream, scratch) for
# its lifetime and hands it back when it dies, so short-lived front-ends reuse warm scratch instead of paying
# vcs_create + cudaMalloc + cudaFree every time.
_free: dict[int, list] = {}


def acquire_context(device: int = 0) -> _capi.Context:
    pool = _free.setdefault(device, [])
    while pool:
        ctx = pool.pop()
        if ctx.h is not None:
            ctx.use_own_stream()
            return ctx
    return _capi.Context(device)


def release_context(ctx) -> None:
    if ctx is not None and getattr(ctx, "h", None) is not None:
        pool = _free.setdefault(ctx.device, [])
        if len(pool) < 4:
            pool.append(ctx)
        else:
            ctx.close()


def close_all():
    for c in _contexts.values():
        c.close()
    _contexts.clear()
    for pool in _free.values():
        for c in pool:
            c.close()
    _free.clear()
