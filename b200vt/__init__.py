"""Import alias: the package directory is `videotuna-dev_b200/` (not a valid Python identifier), so `import b200vt`
loads it from there. Everything lives in that directory; this file only redirects the import machinery."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "videotuna-dev_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
