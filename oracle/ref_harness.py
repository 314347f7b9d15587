"""Loader for the UNMODIFIED reference modules (test infrastructure only).

This file is part of the oracle: only tests/, tests/golden/make_golden.py,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  It never runs on the product path.

It imports /root/reference/code/REC/model/IDNet/hstu.py (HSTU) and
/root/reference/code/REC/evaluator (Collector, Evaluator) without editing them,
by stubbing the four logging-only modules that are not installed in this image
(colorlog, colorama, tensorboardX, pytz) and starting a 1-rank gloo group
(hstu.py:555 calls torch.distributed.get_rank(); basemodel.py:15 calls
get_world_size()).  /root/reference only exists in the build container; on the GPU box
the verbatim copy under oracle/_ref/ (oracle/build_ref.py, git-ignored) is used, and if
that is absent too `available()` is False and callers fall back to the committed golden
fixtures / the restated oracle.
"""
import contextlib
import io
import logging
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_VENDORED = os.path.join(_HERE, "_ref", "code")          # written by oracle/build_ref.py (git-ignored, travels to the GPU box)


def _pick():
    env = os.environ.get("B200REC_REFERENCE")
    for cand in ([env] if env else []) + ["/root/reference/code", _VENDORED]:
        if cand and os.path.isfile(os.path.join(cand, "REC", "model", "IDNet", "hstu.py")):
            return cand
    return "/root/reference/code"


REFERENCE_CODE = _pick()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_CODE, "REC", "model", "IDNet", "hstu.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return
    try:
        __import__(name)
        return
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


_loaded = {}


def ensure_process_group(port: int = 29541):
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(port))
        dist.init_process_group("gloo", rank=0, world_size=1)


def load():
    """Returns (HSTU, Collector, Evaluator) classes of the reference."""
    if _loaded:
        return _loaded["HSTU"], _loaded["Collector"], _loaded["Evaluator"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_CODE)
    import pandas  # noqa: F401  (must precede the pytz stub)
    _stub("colorlog", ColoredFormatter=logging.Formatter)
    _stub("colorama", init=lambda **k: None)
    _stub("tensorboardX", SummaryWriter=object)
    _stub("pytz", utc=None, timezone=lambda n: None)
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)
    ensure_process_group()
    with contextlib.redirect_stdout(io.StringIO()):
        from REC.model.IDNet.hstu import HSTU
        from REC.evaluator import Collector, Evaluator
    _loaded.update(HSTU=HSTU, Collector=Collector, Evaluator=Evaluator)
    return HSTU, Collector, Evaluator


class RefConfig(dict):
    """Mimics REC.config.Config lookups (configurator.py:142-153): a missing key
    reads as None; .get(k, d) returns d when the stored value is None."""

    def __getitem__(self, k):
        return dict.get(self, k, None)

    def get(self, k, d=None):
        v = dict.get(self, k, None)
        return d if v is None else v

    def __contains__(self, k):
        return dict.get(self, k, None) is not None


class RefDataload:
    def __init__(self, item_num, category_counts=None, category_to_int=None):
        self.item_num = item_num
        self.category_counts = category_counts or {}
        self.category_to_int = category_to_int or {}


def build_reference_model(cfg: dict, item_num: int, category_counts=None, category_to_int=None, seed=2020):
    import torch
    HSTU, _, _ = load()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model = HSTU(RefConfig(cfg), RefDataload(item_num, category_counts, category_to_int))
    return model
