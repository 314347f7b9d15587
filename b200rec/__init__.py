"""b200rec — B200-native HSTU multi-head train/eval hot path.

Import shim: the package sources live in
`multi-head-recommendation-with-human-priors_b200/` (not an importable name), so
this module points its `__path__` there.  `import b200rec.hstu` etc. resolve to
files in that directory.
"""
from pathlib import Path as _Path

_SRC = _Path(__file__).resolve().parent.parent / "multi-head-recommendation-with-human-priors_b200"
__path__ = [str(_SRC)]
__version__ = "0.1.0"
