"""Nested dict of tensors with tensor-style indexing: the subset of the reference TensorDict
(habitat-baselines/habitat_baselines/common/tensor_dict.py:57-378) that the trainer, storage and
updater rely on: string keys address children, any other index is applied to every leaf."""
from __future__ import annotations

from typing import Callable, Dict, Union

import numpy as np
import torch

TensorLike = Union[torch.Tensor, np.ndarray]


class TensorDict(dict):
    @classmethod
    def from_tree(cls, tree: Dict) -> "TensorDict":
        res = cls()
        for k, v in tree.items():
            if isinstance(v, dict):
                res[k] = cls.from_tree(v)
            else:
                res[k] = torch.as_tensor(v)
        return res

    def to_tree(self) -> Dict:
        return {k: (v.to_tree() if isinstance(v, TensorDict) else v) for k, v in self.items()}

    def __getitem__(self, index):
        if isinstance(index, str):
            return super().__getitem__(index)
        return TensorDict((k, v[index]) for k, v in self.items())

    def set(self, index, value, strict: bool = True):
        if isinstance(index, str):
            if isinstance(value, dict) and not isinstance(value, TensorDict):
                value = TensorDict.from_tree(value)
            super().__setitem__(index, value)
            return
        if strict and set(self.keys()) != set(value.keys()):
            raise KeyError(f"Keys don't match: Dest={list(self.keys())} Source={list(value.keys())}")
        for k in self.keys():
            if k not in value:
                if strict:
                    raise KeyError(f"Key {k} not in new value dictionary")
                continue
            v = value[k]
            dst = super().__getitem__(k)
            if isinstance(v, (TensorDict, dict)):
                dst.set(index, v, strict=strict)
            else:
                dst[index].copy_(torch.as_tensor(v))

    def __setitem__(self, index, value):
        self.set(index, value)

    def map_func(self, func: Callable, src, dst=None, needs_grad=False):
        return self.map(func)

    def map(self, func: Callable[[torch.Tensor], torch.Tensor]) -> "TensorDict":
        return TensorDict((k, v.map(func) if isinstance(v, TensorDict) else func(v)) for k, v in self.items())

    def map_in_place(self, func: Callable[[torch.Tensor], torch.Tensor]) -> "TensorDict":
        for k, v in list(self.items()):
            if isinstance(v, TensorDict):
                v.map_in_place(func)
            else:
                super().__setitem__(k, func(v))
        return self

    def slice_keys(self, *keys) -> "TensorDict":
        res = TensorDict()
        for ks in keys:
            for k in ([ks] if isinstance(ks, str) else ks):
                res.set(k, super().__getitem__(k))
        return res

    def __deepcopy__(self, memo):
        return TensorDict((k, v.clone() if torch.is_tensor(v) else v.__deepcopy__(memo)) for k, v in self.items())
