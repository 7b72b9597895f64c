"""Parameter: a Tensor that shares its source buffer and registers with Modules (reference nn/parameter.py:4-17)."""
from ..core import Tensor


def _note_rebind():
    from . import _plans
    _plans.note_structure_change()


class Parameter(Tensor):

    def __init__(self, data: Tensor, requires_grad: bool = True) -> None:
        super().__init__(data=data.data, dtype=data.dtype, device=data.device, copy=False, requires_grad=requires_grad)

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name == "data":  # re-bound storage (p.data = w, .to(device)): recorded inference plans must re-check their pointers
            _note_rebind()

    def __repr__(self) -> str:
        return "Parameter : \n{}".format(self.data) + (",\ndevice={}".format(self.device) if self.device.device != "cpu" else "")
