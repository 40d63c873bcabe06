"""CPU: libfesr.so builds, loads and exports exactly what include/fesr.h declares (no compute)."""
import os
import re

import pytest

from conftest import ROOT
from fesr_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from fesr_b200 import build
        build.build()
    return _lib.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "fesr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fesr_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_typed(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in fesr.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_dims(lib):
    assert lib.fesr_version() == 100
    d = _lib.model_dims(_lib.KERNELNN, 43, 4, 4, 5)
    assert (d.wp, d.k1, d.kt, d.passes, d.kp, d.k1p, d.zk_main, d.zk) == (48, 44, 11, 1, 48, 44, 2112, 2176)
    t = _lib.model_dims(_lib.TEECNET, 43, 4, 4, 5)
    assert (t.wp, t.k1, t.kt, t.passes, t.kp, t.k1p, t.zk) == (48, 129, 11, 3, 144, 132, 6400)
    d48 = _lib.model_dims(_lib.KERNELNN, 48, 4, 4, 5)
    assert (d48.wp, d48.k1, d48.kt, d48.passes, d48.zk) == (48, 49, 13, 1, 2560)
    t48 = _lib.model_dims(_lib.TEECNET, 48, 4, 4, 5)
    assert t48.wp == 64


def test_bad_arguments_are_rejected_with_a_message(lib):
    with pytest.raises(_lib.FesrError, match="width"):
        _lib.model_dims(_lib.KERNELNN, 100, 4, 4, 5)
    with pytest.raises(_lib.FesrError, match="kind"):
        _lib.model_dims(7, 43, 4, 4, 5)


def test_models_refuse_cpu_tensors():
    import torch
    from fesr_b200.models.model import KernelNN
    m = KernelNN(8, 8, 2, in_width=4, out_width=4)
    with pytest.raises(_lib.FesrError, match="no CPU"):
        m(torch.zeros(3, 4), torch.zeros(2, 0, dtype=torch.long), torch.zeros(0))


def test_state_dict_keys_match_reference_checkpoints(golden):
    from conftest import state_dict_from
    from fesr_b200.models.model import KernelNN, TEECNet
    k = KernelNN(16, 16, 3, in_width=4, out_width=4)
    k.load_state_dict(state_dict_from(golden, "kernelnn_w16"), strict=True)
    t = TEECNet(4, 12, 4, num_layers=2, retrieve_weight=False)
    t.load_state_dict(state_dict_from(golden, "teecnet_w12"), strict=True)
