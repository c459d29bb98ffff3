"""Drop-in shim: put this directory in front of the reference on sys.path and
`import defects` resolves to the B200 implementation (cetkmc.defects)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
import cetkmc  # noqa: E402,F401
from cetkmc.defects import *  # noqa: E402,F401,F403
from cetkmc import defects as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
