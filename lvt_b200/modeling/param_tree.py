"""nn.Module trees whose parameters are VIEWS of an engine's flat fp32 master buffer (and whose .grad are
views of the flat gradient), named exactly like the reference's state_dict keys."""
import torch
from torch import nn


class ParamTree(nn.Module):
    """Container that grows child containers along dotted names: add('a.b.0.weight', tensor)."""

    def _child(self, name):
        if name not in self._modules:
            self.add_module(name, ParamTree())
        return self._modules[name]

    def add_param(self, dotted, value, grad=None, requires_grad=True):
        node = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            node = node._child(p)
        param = nn.Parameter(value, requires_grad=requires_grad)
        if grad is not None and requires_grad:
            param.grad = grad
        node.register_parameter(parts[-1], param)
        return param

    def add_buffer(self, dotted, value, persistent=True):
        node = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            node = node._child(p)
        node.register_buffer(parts[-1], value, persistent=persistent)


def attach_store(tree: ParamTree, store, prefix="", strip=""):
    """Register every parameter of `store` whose name starts with `prefix` under its name minus `strip`."""
    for name in store.shapes:
        if name.startswith(prefix):
            tree.add_param(name[len(strip):], store.p[name], store.g[name])
    return tree
