"""Importable alias of the package directory `vlm-bridge-for-image-captioning_b200/`.

The repo layout names the package after the reference repository; that name contains hyphens and
cannot be written in an `import` statement, so this shim re-roots `vlm_bridge_b200` onto it.
"""
import os as _os

_real = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "vlm-bridge-for-image-captioning_b200",
)
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
