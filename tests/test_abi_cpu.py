"""CPU-side checks of the C ABI: the library builds for sm_100a, loads, and exports every symbol
declared in include/b200rec.h; the ctypes table matches the header; the product path fails loudly
without CUDA (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200rec.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rec_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    from b200rec import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        subprocess.check_call(["make", "-j8", "-C", os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc")])
    return _lib.LIB_PATH


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_covers_header(libpath):
    from b200rec import _lib
    declared = set(_declared())
    bound = set(_lib.EXPORTS)
    assert declared <= bound | {"b200rec_last_error"}, sorted(declared - bound)
    lib = ctypes.CDLL(libpath)
    assert lib.b200rec_version() >= 100


def test_no_cpu_fallback():
    from b200rec import synth, _lib
    from b200rec.hstu import HSTU
    cfg = synth.make_config("A", item_num=200, train_batch_size=4, num_negatives=16)
    model = HSTU(cfg, synth.make_dataload(cfg), compute_dtype=torch.float32)
    batch = synth.make_train_batch(cfg, seed=0)
    with pytest.raises(_lib.B200RecError):
        model(batch)
    with pytest.raises(_lib.B200RecError):
        model.predict(torch.ones(2, 20, dtype=torch.int64), None, torch.randn(200, 64), None, None)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multi-head-recommendation-with-human-priors_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_unsupported_configs_raise():
    from b200rec import synth
    from b200rec.hstu import HSTU
    for over in (dict(prior_switch="in_out", prior_switch_loss_weight=1.0), dict(pos_sample_mix_ratio=0.1),
                 dict(head_interaction="hierarchical", cat_bottleneck=True, cat_bottleneck_dim=12)):   # bf16: 16-byte TMA pitch
        cfg = synth.make_config("D", item_num=200, **over)
        with pytest.raises(NotImplementedError):
            HSTU(cfg, synth.make_dataload(cfg))
    cfg = synth.make_config("D", item_num=200, prior_switch="in", prior_switch_loss_weight=0.5)
    m = HSTU(cfg, synth.make_dataload(cfg))            # prior-switch aux heads ('in') are built
    assert len(m.aux_cat_head) == cfg["num_prior_head"] and "aux_cat_head.0.weight" in m.state_dict()
