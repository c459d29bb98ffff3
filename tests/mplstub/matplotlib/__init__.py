"""Test infrastructure: a do-nothing matplotlib for running the reference's plotting callers
(main.py, plot_cet.py, lattice_init.py import it at module level; matplotlib is not installed in
this image).  Every attribute is a callable that returns another such object."""


class _Any:
    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Any()

    def __iter__(self):
        return iter((_Any(), _Any()))

    def __getitem__(self, key):
        return _Any()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __len__(self):
        return 2

    def __float__(self):
        return 0.0

    def __bool__(self):
        return True


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return _Any()
