"""Import shim: the product package lives in the directory ``video-filler_b200/`` (the name the
project layout prescribes, which is not a valid Python identifier).  ``import video_filler_b200``
resolves to that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "video-filler_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
