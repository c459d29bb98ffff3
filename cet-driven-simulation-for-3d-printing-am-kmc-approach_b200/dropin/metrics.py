"""Drop-in shim: put this directory in front of the reference on sys.path and
`import metrics` resolves to the B200 implementation (cetkmc.metrics)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
import cetkmc  # noqa: E402,F401
from cetkmc.metrics import *  # noqa: E402,F401,F403
from cetkmc import metrics as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
