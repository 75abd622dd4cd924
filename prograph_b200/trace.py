"""Tracing of the graph build: NVTX ranges around every phase (pack / boot / sweep / exchange /
merge / gather / finalize / d2h) and, when enabled, CUDA-event timing of the same ranges.

The reference has no tracing; this is the hook nsys / ncu --nvtx see, and what
bench.py reads to report where a step's time goes.  Events are recorded on torch's current
stream -- the stream every library call is launched on -- and only resolved when `phases()` is
called, so tracing never synchronises the build.
"""
import contextlib

import torch

_timing = False
_events = []          # (name, start event, end event)


def enable_timing(on=True):
    """Record a CUDA event pair per phase (read them back with phases())."""
    global _timing
    _timing = bool(on)
    if not on:
        _events.clear()


@contextlib.contextmanager
def phase(name):
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.nvtx.range_push("pg:" + name)
    rec = None
    if cuda and _timing:
        rec = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        rec[0].record()
    try:
        yield
    finally:
        if rec is not None:
            rec[1].record()
            _events.append((name, rec[0], rec[1]))
        if cuda:
            torch.cuda.nvtx.range_pop()


def phases(reset=True):
    """{phase name: total milliseconds} over the recorded ranges (synchronises)."""
    out = {}
    for name, a, b in _events:
        b.synchronize()
        out[name] = out.get(name, 0.0) + a.elapsed_time(b)
    if reset:
        _events.clear()
    return out
