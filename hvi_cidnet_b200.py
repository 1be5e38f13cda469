"""Import shim: `import hvi_cidnet_b200` resolves to the package directory
`hvi-cidnet_b200/` (a hyphen is not importable, the directory name is fixed by
the repo layout)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "hvi-cidnet_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
