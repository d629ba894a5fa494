"""Import alias for the product package.

The product lives in ``motion-generation-using-quadratic-programs_b200/`` (the directory name the
project layout fixes); hyphens make that name un-importable, so this alias package points its
``__path__`` there: ``import mgqp_b200.engine`` loads
``motion-generation-using-quadratic-programs_b200/engine.py``.
"""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "motion-generation-using-quadratic-programs_b200")
__path__.insert(0, _impl)

from ._api import *  # noqa: F401,F403,E402
from ._api import __all__  # noqa: E402
