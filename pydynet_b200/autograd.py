"""Global grad-mode switch and the ``no_grad`` / ``enable_grad`` context-manager-decorators.

Same surface as reference pydynet/autograd.py:3-50 (a process-global flag; ``Module.train(mode)`` flips it
too, reference nn/modules/module.py:45-47).
"""
import functools

grad_enable = True


def is_grad_enable() -> bool:
    return grad_enable


def set_grad_enabled(mode: bool) -> None:
    global grad_enable
    grad_enable = bool(mode)


class _GradMode:
    _mode = True

    def __enter__(self):
        self.prev = is_grad_enable()
        set_grad_enabled(self._mode)

    def __exit__(self, *exc):
        set_grad_enabled(self.prev)

    def __call__(self, func):
        cls = type(self)

        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            with cls():
                return func(*args, **kwargs)

        return wrapper


class no_grad(_GradMode):
    _mode = False


class enable_grad(_GradMode):
    _mode = True
