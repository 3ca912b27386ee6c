"""CPU-side checks of the boundary: the library loads, exports every symbol the
header declares, and refuses to run without CUDA tensors (no fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from mimrl_b200 import _lib
    return _lib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mimrl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mimrl_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib.lib, s), f"{s} declared in include/mimrl_b200.h but not exported"
        assert s in lib.EXPORTS, f"{s} has no ctypes signature in mimrl_b200/_lib.py"
    assert sorted(lib.EXPORTS) == syms


def test_version_and_error_string(lib):
    assert lib.lib.mimrl_version() == 1
    assert isinstance(lib.lib.mimrl_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # validation happens before any launch, so it is testable on a CPU-only host
    rc = lib.lib.mimrl_sep_row_stats(None, None, 0, 0, 0, 0, 0, 0, None, None, None, None, None, 0, None)
    assert rc != 0 and b"empty" in lib.lib.mimrl_last_error()
    rc = lib.lib.mimrl_knn_search(None, 10, 4, None, 4, 9, 1.0, 1, None, None, None, None, 0, None)
    assert rc != 0 and b"n_neighbors" in lib.lib.mimrl_last_error()
    assert lib.lib.mimrl_bound_weight_family(8, None, None, None) != 0          # interpolate has no single-sweep form


def test_no_cpu_fallback(lib):
    from mimrl_b200.vmi import separable_bound, infonce_lower_bound
    x = torch.randn(8, 16)
    with pytest.raises(lib.MimrlError):
        separable_bound(x, x, "infonce")
    with pytest.raises(lib.MimrlError):
        infonce_lower_bound(torch.randn(8, 8))


def test_unknown_types_raise_like_reference():
    from mimrl_b200.vmi import BaselineModel, CriticModel
    from mimrl_b200.model import MLP_For_CMI
    with pytest.raises(NotImplementedError):
        CriticModel("bilinear", 8, 8)
    with pytest.raises(NotImplementedError):
        BaselineModel("learned", 8)
    with pytest.raises(NotImplementedError):
        MLP_For_CMI(8, 8, 2, 2, "relu", "tanh")


def test_state_dict_names_match_reference():
    from mimrl_b200.model import VCMIEstimator, VMIEstimator
    from oracle import params as P
    est = VMIEstimator("separate", "unnormalized", "tuba", 16, 32, 16, 2, "relu", 0, 1)
    want = set(P.vmi_state_dict(P.vmi_params(0, "separate", "unnormalized", 16, 32, 16, 2)))
    assert set(est.state_dict()) == want
    est = VMIEstimator("concat", "constant", "nwj", 16, 32, 16, 2, "relu", 0, 1)
    assert set(est.state_dict()) == set(P.vmi_state_dict(P.vmi_params(0, "concat", "constant", 16, 32, 16, 2)))
    c = VCMIEstimator(16, 32, 2, "relu", 2, 1.0)
    assert set(c.state_dict()) == set(P.vcmi_state_dict(P.vcmi_params(0, 16, 32)))
    # biases start at zero (VMI.py:47-51)
    assert all(float(v.abs().max()) == 0.0 for k, v in est.state_dict().items() if k.endswith("bias"))


def test_knn_size_functions_without_gpu(lib):
    """Host-only planning of the k-NN entry points: which pools have a fit (tensor-core route: width > 15 and at least
    2048 rows), and the width-1 workspace carries the sorted pool (four N-entry arrays + the radix-sort scratch)."""
    N = 1 << 20
    assert lib.lib.mimrl_knn_fit_bytes(N, 128) >= 2 * N * 128 * 2 + N * 4          # fp16 hi / lo planes + norms
    assert lib.lib.mimrl_knn_fit_bytes(5000, 40) > 0
    assert lib.lib.mimrl_knn_fit_bytes(N, 1) == 0 and lib.lib.mimrl_knn_fit_bytes(1000, 128) == 0 and lib.lib.mimrl_knn_fit_bytes(0, 128) == 0
    assert lib.lib.mimrl_knn_workspace_bytes(N, 4096, 1, 2) >= 4 * N * 4 + N * 4 > lib.lib.mimrl_knn_workspace_bytes(N, 4096, 2, 2)
