"""Import shim: the package sources live in ``dfd-clip_b200/`` (the name the project layout fixes, which
is not a legal Python identifier). ``import dfdclip_b200`` resolves every submodule from that directory."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dfd-clip_b200")
__path__ = [_PKG_DIR]

with open(_os.path.join(_PKG_DIR, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
