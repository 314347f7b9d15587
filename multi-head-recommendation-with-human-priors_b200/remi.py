"""`REC.model.IDNet.remi` counterpart: the reference resolves a model as module `<name.lower()>`, attribute `<name>`
(REC/utils/utils.py:38-57), so `REMI` is importable from a module of its own.  The class lives in comirec.py."""
from .comirec import REMI  # noqa: F401
