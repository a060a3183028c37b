"""wlsqm_b200.utils.ptrwrap -- the reference's ``PointerWrapper`` (``wlsqm/utils/ptrwrap.pyx``): an object that carries a
C pointer through Python (``ExpertSolver.manager_pw``, ``expert.pyx:257-260``).  The drop-in's native handle is a
``ctypes.c_void_p`` to a ``wlsqm_solver``; this class gives it the reference's shape."""

__all__ = ["PointerWrapper"]


class PointerWrapper:
    """holds ``ptr`` (an address as int, or None)"""
    __slots__ = ("ptr",)

    def __init__(self, ptr=None):
        self.ptr = ptr

    def set_ptr(self, ptr):
        self.ptr = ptr
