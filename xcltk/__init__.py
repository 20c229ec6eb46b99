"""`xcltk` import name for the B200 build: `from xcltk.rdr.fc.main import fc_wrapper`,
`from xcltk.baf.fc.main import afc_wrapper` and `xcltk.xcltk:main` (the reference's public entry points,
xcltk/rdr/fc/main.py:142, xcltk/baf/fc/main.py:32, xcltk/xcltk.py:40) resolve to the modules of `xcltk_b200`.

Nothing lives here: the package's search path is xcltk_b200's, so every submodule import finds the
implementation under its reference name.
"""

import xcltk_b200 as _impl
from xcltk_b200.config import APP, VERSION  # noqa: F401

__path__ = list(_impl.__path__)
__version__ = VERSION
