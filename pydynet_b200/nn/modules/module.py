"""Module system (reference nn/modules/module.py:8-143): flat dotted ``_parameters`` registry filled at attribute
assignment, ``train(mode)`` that also flips the global grad switch (reference module.py:45-47), device migration of
Parameter attributes."""
from collections import OrderedDict

from ..parameter import Parameter
from ...autograd import set_grad_enabled
from ...core import Tensor
from ...cuda import Device, current_device
from .. import _plans


class Module:

    def __init__(self) -> None:
        self._train = True
        self.device = Device("cpu")
        self._parameters = OrderedDict()

    def __call__(self, *x):
        # eval-mode calls on a cuda device may be served by a fused inference plan (nn/_plans.py) whose results are those of
        # forward(); anything a plan does not cover falls through to the eager forward like the reference (module.py:16-17)
        plan = self.__dict__.get("_pdn_plan")
        if plan is None:
            plan = _plans.attach(self)
        if plan is not False:
            out = plan(x)
            if out is not NotImplemented:
                return out
        return self.forward(*x)

    def __setattr__(self, name: str, value) -> None:
        object.__setattr__(self, name, value)
        if isinstance(value, Parameter):
            self._parameters[name] = value
            _plans.note_structure_change()
        elif isinstance(value, Module):
            for key, p in value._parameters.items():
                self._parameters[name + "." + key] = p
            _plans.note_structure_change()

    def _children(self):
        return [(k, v) for k, v in self.__dict__.items() if isinstance(v, Module)]

    def __repr__(self) -> str:
        return "{}(\n{}\n)".format(self.__class__.__name__, "\n".join("{:>10} : {}".format(k, m) for k, m in self._children()))

    def parameters(self):
        for p in self._parameters.values():
            if p.requires_grad:
                yield p

    def train(self, mode: bool = True):
        set_grad_enabled(mode)
        self.set_module_state(mode)

    def eval(self):
        return self.train(False)

    def set_module_state(self, mode: bool):
        self._train = mode
        for _, m in self._children():
            m.set_module_state(mode)

    def forward(self, x: Tensor) -> Tensor:
        raise NotImplementedError

    def to(self, device):
        device = Device(device)
        if self.device != device:
            self.move(device)
        return self

    def move(self, device):
        self.device = device
        for v in list(self.__dict__.values()):
            if isinstance(v, Module):
                v.move(device)
            elif isinstance(v, Parameter):
                v.to(device)

    def cuda(self):
        return self.to(current_device())

    def cpu(self):
        return self.to('cpu')


class Sequential(Module):

    def __init__(self, *args) -> None:
        super().__init__()
        self.module_list = []
        items = args[0].items() if len(args) == 1 and isinstance(args[0], OrderedDict) else ((str(i), m) for i, m in enumerate(args))
        for name, module in items:
            setattr(self, name, module)
            self.module_list.append(module)

    def forward(self, x):
        for module in self.module_list:
            x = module(x)
        return x

    def __len__(self):
        return len(self.module_list)


class ModuleList(Module):

    def __init__(self, module_list: list) -> None:
        super().__init__()
        self.module_list = module_list
        for idx, module in enumerate(module_list):
            setattr(self, str(idx), module)

    def __getitem__(self, index):
        return self.module_list[index]

    def __len__(self):
        return len(self.module_list)

    def __iter__(self):
        return iter(self.module_list)

    def append(self, module):
        self.module_list.append(module)
        setattr(self, str(len(self.module_list) - 1), module)

    def index(self, module):
        return self.module_list.index(module)
