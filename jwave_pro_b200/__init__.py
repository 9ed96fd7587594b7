"""Import shim: the package directory is named `jwave-pro_b200/` (not a Python identifier), so
`import jwave_pro_b200` resolves its submodules from that directory."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "jwave-pro_b200"))

from ._pkg import *  # noqa: F401,F403,E402
from ._pkg import __all__  # noqa: E402
