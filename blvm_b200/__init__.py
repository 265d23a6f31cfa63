"""Import alias: the package lives in `benchmarking-lvms_b200/` (a directory name Python cannot import because of the
hyphen); `import blvm_b200` loads it from there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "benchmarking-lvms_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"), globals())
del _os, _f, _real
