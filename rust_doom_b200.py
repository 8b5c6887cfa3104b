"""Import shim: the package directory is `rust-doom_b200/` (hyphenated, as the project is named), which
Python cannot import by that name.  `import rust_doom_b200` lands here and is replaced by the package."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "rust-doom_b200")
_spec = _ilu.spec_from_file_location("rust_doom_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["rust_doom_b200"] = _mod
_spec.loader.exec_module(_mod)
