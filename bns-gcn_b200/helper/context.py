"""Per-rank singletons: ``ctx.buffer``, ``ctx.reducer`` (helper/context.py:4-5 of the reference) and the
communicator.  The reference keeps them as module globals (one process per rank); here they are
thread-local so that P ranks can also live as threads of one process (tests, smoke, 1-GPU emulation)."""
import threading

_tls = threading.local()


class _Proxy:
    """Attribute access is forwarded to the calling thread's instance (created on first use)."""

    def __init__(self, name, factory):
        object.__setattr__(self, "_name", name)
        object.__setattr__(self, "_factory", factory)

    def _get(self):
        inst = getattr(_tls, self._name, None)
        if inst is None:
            inst = self._factory()
            setattr(_tls, self._name, inst)
        return inst

    def __getattr__(self, item):
        return getattr(self._get(), item)

    def __setattr__(self, key, value):
        setattr(self._get(), key, value)


def _make_buffer():
    from .feature_buffer import Buffer
    return Buffer()


def _make_reducer():
    from .reducer import Reducer
    return Reducer()


buffer = _Proxy("buffer", _make_buffer)
reducer = _Proxy("reducer", _make_reducer)


def set_comm(c):
    _tls.comm = c


def comm():
    c = getattr(_tls, "comm", None)
    if c is None:
        import torch.distributed as dist
        from .comm import DistComm, SoloComm
        c = DistComm() if dist.is_available() and dist.is_initialized() else SoloComm()
        _tls.comm = c
    return c


def reset():
    for k in ("buffer", "reducer", "comm"):
        if hasattr(_tls, k):
            delattr(_tls, k)
