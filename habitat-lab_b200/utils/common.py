"""Host-side observation batching (habitat-baselines/habitat_baselines/utils/common.py:191-330): per-environment
observation dicts -> one TensorDict of [N, ...] tensors on the learner's device.

Data movement only (no arithmetic): every sensor is staged once in a reusable pinned host buffer and sent with one
asynchronous H2D copy per sensor, instead of the reference's N per-env copies per sensor.  Observations that are
already torch tensors on the target device (a GPU simulator, the synthetic env) are stacked in place."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from ..common.tensor_dict import TensorDict


class ObservationBatchingCache:
    """Reusable pinned staging buffers keyed by (sensor, shape, dtype) -- the role of `_ObservationBatchingCache`."""

    def __init__(self):
        self._pool: Dict[tuple, torch.Tensor] = {}
        self._in_flight: Dict[tuple, "torch.cuda.Event"] = {}

    def get(self, name: str, shape, dtype: torch.dtype, pin: bool) -> torch.Tensor:
        """The staging buffer for `name`; blocks until the previous asynchronous H2D copy out of it has finished
        (a second batch_obs with the same cache -- e.g. a double-buffered sampler -- must not overwrite host memory
        the DMA engine is still reading)."""
        key = (name, tuple(shape), dtype, pin)
        buf = self._pool.get(key)
        if buf is None:
            buf = torch.empty(tuple(shape), dtype=dtype, pin_memory=pin)
            self._pool[key] = buf
        ev = self._in_flight.pop(key, None)
        if ev is not None:
            ev.synchronize()
        return buf

    def mark_in_flight(self, name: str, shape, dtype: torch.dtype, pin: bool) -> None:
        ev = torch.cuda.Event()
        ev.record()
        self._in_flight[(name, tuple(shape), dtype, pin)] = ev


def batch_obs(observations: List[dict], device: Optional[torch.device] = None,
              cache: Optional[ObservationBatchingCache] = None) -> TensorDict:
    """observations: one dict per environment (numpy arrays, scalars or tensors, possibly nested).  Returns a
    TensorDict whose leaves have a leading env dimension.  uint8 stays uint8 and float64 becomes float32, as in the
    reference (common.py:262-330)."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    cache = cache if cache is not None else ObservationBatchingCache()
    pin = device.type == "cuda" and torch.cuda.is_available()

    def build(items: List, name: str):
        first = items[0]
        if isinstance(first, dict):
            return {k: build([it[k] for it in items], f"{name}/{k}") for k in first}
        if torch.is_tensor(first):
            if first.device == device:
                return torch.stack(list(items), 0)
            items = [t.cpu().numpy() for t in items]
            first = items[0]
        arr0 = np.asarray(first)
        dtype = torch.from_numpy(arr0.reshape(-1)[:0].copy()).dtype if arr0.dtype != np.float64 else torch.float32
        buf = cache.get(name, (len(items),) + arr0.shape, dtype, pin)
        view = buf.numpy()
        for i, it in enumerate(items):
            view[i] = np.asarray(it)
        if device.type != "cuda":
            return buf.clone()
        out = buf.to(device, non_blocking=True)
        if pin:
            cache.mark_in_flight(name, (len(items),) + arr0.shape, dtype, pin)
        return out

    return TensorDict.from_tree(build(list(observations), "obs"))
