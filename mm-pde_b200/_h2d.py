"""Host -> device staging of the per-step input slices (the reference does ``data.to(device)`` at every use:
data_creator_2d.py:157-267, train_helper_2d.py:121).

A pageable ``.to(device)`` blocks the host until the stream has drained, once per call -- and the same slice is moved
up to five times per step (two create_graph calls, interpolate_pred, the loss).  Here every distinct host tensor is
copied ONCE, through a small ring of pinned buffers with ``non_blocking=True``, and the device copy is remembered
while the host tensor is alive and unmodified; the host keeps queueing kernels while the copy engine works."""
import weakref

import torch

_RING = 4
_rings = {}          # (numel, dtype) -> [slot index, [pinned buffers], [events]]
_cache = {}          # id(host tensor) -> (weakref, version, device, device tensor)


def _pinned_slot(numel, dtype):
    key = (numel, dtype)
    ring = _rings.get(key)
    if ring is None:
        ring = _rings[key] = [0, [torch.empty(numel, dtype=dtype).pin_memory() for _ in range(_RING)], [None] * _RING]
    i = ring[0]
    ring[0] = (i + 1) % _RING
    if ring[2][i] is not None:
        ring[2][i].synchronize()          # the copy that last used this slot has long finished; never blocks in practice
    return ring, i


def to_device(t, device):
    device = torch.device(device)
    if not isinstance(t, torch.Tensor) or t.device == device:
        return t
    if device.type != "cuda" or t.is_cuda:
        return t.to(device)
    hit = _cache.get(id(t))
    if hit is not None and hit[0]() is t and hit[1] == t._version and hit[2] == device:
        return hit[3]
    src = t.contiguous()
    ring, i = _pinned_slot(src.numel(), src.dtype)
    stage = ring[1][i]
    stage.copy_(src.reshape(-1))
    out = torch.empty(src.shape, dtype=src.dtype, device=device)
    out.reshape(-1).copy_(stage, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    ring[2][i] = ev
    if len(_cache) > 64:
        for k in [k for k, v in _cache.items() if v[0]() is None]:
            del _cache[k]
        if len(_cache) > 64:
            _cache.clear()
    _cache[id(t)] = (weakref.ref(t), t._version, device, out)
    return out
