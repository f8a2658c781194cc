"""Notifier — the observer pattern Cherab uses for cache control between disconnected objects
(cherab/core/utility/notify.py:33-162): callbacks without arguments, held by weak reference so that a registered
object can still be garbage collected; `notify()` calls every live callback and drops the dead ones."""
from types import BuiltinMethodType, MethodType
from weakref import ref


class _Notifier:

    def __init__(self):
        self._callbacks_refs = []

    @staticmethod
    def _is_method(callback):
        return isinstance(callback, (MethodType, BuiltinMethodType))

    def add(self, callback):
        """Register a callback (function or bound method); registering twice has no effect (notify.py:62-76)."""
        if self.is_present(callback):
            return
        if self._is_method(callback):
            self._callbacks_refs.append((ref(callback.__self__), callback.__name__))
        else:
            self._callbacks_refs.append(ref(callback))

    def remove(self, callback):
        for reference in list(self._callbacks_refs):
            if self._matches(reference, callback):
                self._callbacks_refs.remove(reference)
                break

    def is_present(self, callback):
        return any(self._matches(reference, callback) for reference in self._callbacks_refs)

    def _matches(self, reference, callback):
        if isinstance(reference, tuple):
            return self._is_method(callback) and reference[0]() is callback.__self__ and reference[1] == callback.__name__
        return (not self._is_method(callback)) and reference() is callback

    def notify(self):
        """Call every registered callback; references whose object died are purged (notify.py:107-137)."""
        dead = []
        for reference in list(self._callbacks_refs):
            if isinstance(reference, tuple):
                instance = reference[0]()
                if instance is None:
                    dead.append(reference)
                    continue
                getattr(instance, reference[1])()
            else:
                callback = reference()
                if callback is None:
                    dead.append(reference)
                    continue
                callback()
        for reference in dead:
            if reference in self._callbacks_refs:
                self._callbacks_refs.remove(reference)


# Where Cherab itself is importable its own class is the one in use (host plumbing to reuse as it is); the class
# above is its stand-in for environments without Cherab / Raysect, like the build image.
try:
    from cherab.core.utility import Notifier            # noqa: F401
except Exception:                                       # ImportError, or a Cherab that fails to load without Raysect
    Notifier = _Notifier
