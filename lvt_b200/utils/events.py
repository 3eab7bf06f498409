"""Minimal EventStorage (reference vidgen/utils/events.py:16-25,233-375): the model classes only need
`get_event_storage().iter` inside a `with EventStorage(start_iter):` block, plus put_scalar/put_image."""
from collections import defaultdict

_CURRENT_STORAGE_STACK = []


def get_event_storage():
    assert len(_CURRENT_STORAGE_STACK), \
        "get_event_storage() has to be called inside a 'with EventStorage(...)' context!"
    return _CURRENT_STORAGE_STACK[-1]


class EventStorage:
    def __init__(self, start_iter=0):
        self._history = defaultdict(list)
        self._latest = {}
        self._images = []
        self._iter = start_iter

    def put_scalar(self, name, value, smoothing_hint=True):
        value = float(value)
        self._history[name].append((value, self._iter))
        self._latest[name] = value

    def put_scalars(self, *, smoothing_hint=True, **kwargs):
        for k, v in kwargs.items():
            self.put_scalar(k, v, smoothing_hint)

    def put_image(self, name, img):
        self._images.append((name, img, self._iter))

    def clear_images(self):
        self._images = []

    def history(self, name):
        return self._history[name]

    def latest(self):
        return self._latest

    def step(self):
        self._iter += 1
        self._latest = {}

    @property
    def iter(self):
        return self._iter

    @property
    def iteration(self):
        return self._iter

    def __enter__(self):
        _CURRENT_STORAGE_STACK.append(self)
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        assert _CURRENT_STORAGE_STACK[-1] == self
        _CURRENT_STORAGE_STACK.pop()
