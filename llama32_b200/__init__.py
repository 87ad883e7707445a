"""Importable alias of the `llama-3.2-multimodal_b200/` package directory (its name is not an identifier).

`import llama32_b200` exposes the drop-in modules; submodules (`llama32_b200.ops`, `.modules`, `._lib`,
`.build`, `.tp`) resolve into `llama-3.2-multimodal_b200/`.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "llama-3.2-multimodal_b200"))

from .modules import *  # noqa: E402,F401,F403
from .modules import __all__  # noqa: E402,F401
