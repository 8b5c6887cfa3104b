"""Import shim: the package directory is `rust-doom_b200/` (hyphen, as the project is named), which
Python cannot import directly.  `import rust_doom_b200` resolves here and behaves as that package."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "rust-doom_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
