"""Import shim: the product package lives in the directory `video-stab_b200/` (the name the
project brief fixes), which is not a valid Python identifier.  `import video_stab_b200`
resolves to that directory."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "video-stab_b200"))
from ._pkg import *  # noqa: F401,F403,E402
from ._pkg import __all__  # noqa: E402
