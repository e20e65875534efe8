"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/fabric_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fabric_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fabric_b200_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    from fabric_b200 import _lib
    assert header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_symbol(built_lib):
    for name in header_symbols():
        assert hasattr(built_lib, name), name
    assert built_lib.fabric_b200_version() >= 100
    assert isinstance(built_lib.fabric_b200_last_error(), bytes)


def test_no_gpu_means_loud_failure(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fabric_b200 import BiDateNet
    m = BiDateNet(13, 2).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 13, 32, 32), torch.zeros(1, 13, 32, 32))
    assert built_lib.fabric_b200_sm_count() < 0       # no device: error code, not a fallback


def test_state_dict_keys_match_reference_spec():
    import torch
    from fabric_b200 import BiDateNet
    from oracle import bidatenet_oracle as O
    m = BiDateNet(13, 2)
    sd = m.state_dict()
    spec = O.state_dict_spec()
    assert list(sd.keys()) == [k for k, _, _ in spec]
    for k, shape, dt in spec:
        assert tuple(sd[k].shape) == tuple(shape) and sd[k].dtype == dt, k
    m.load_state_dict(O.make_state_dict(0))            # lossless load of reference-shaped weights
    import models.bidate_model as ref_path             # the reference's pickle path resolves to the same class
    assert ref_path.BiDateNet is BiDateNet
    import io, pickle  # noqa
    buf = io.BytesIO()
    torch.save(m, buf)                                  # train.py:222 pickles the whole module
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert list(m2.state_dict().keys()) == list(sd.keys())
