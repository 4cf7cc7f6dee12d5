"""The C-ABI library loads on a CPU-only box and exports every symbol include/bsnative.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bsnative.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bsnative.h but not exported"


def test_python_binding_lists_the_same_symbols(built_lib):
    from bootstrapper_b200 import native
    assert sorted(native.EXPORTS) == declared_symbols()
    native.lib()
    assert native.lib().bs_version() >= 100


def test_error_reporting_without_compute(built_lib):
    from bootstrapper_b200 import native
    with pytest.raises(native.BsError) as e:
        native.Plan((10, 10, 10), (0, 5, 5), (0, 1, 1), native.BS_DTYPE_U8)
    assert "positive" in str(e.value)
    with pytest.raises(native.BsError):
        native.Plan((10, 10, 10), (5, 5, 5), (6, 1, 1), native.BS_DTYPE_U8)     # context > block


def test_no_cpu_fallback(built_lib):
    """CPU tensors are refused, the product never routes through the oracle."""
    import torch
    from bootstrapper_b200 import native
    from bootstrapper_b200.post.pipeline import segment_blockwise
    with pytest.raises(native.BsError):
        segment_blockwise(torch.zeros((3, 4, 8, 8), dtype=torch.uint8))
    pkg = os.path.join(ROOT, "bootstrapper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_config_struct_layout(built_lib):
    """the ctypes mirror of bs_ws_config has the size the library was compiled with (checked again at load time)"""
    import ctypes as C
    from bootstrapper_b200 import native
    lib = native.lib()
    assert lib.bs_config_size() == C.sizeof(native.WsConfig)
    cfg = native.WsConfig()
    assert C.sizeof(cfg.bias) == 24 and native.WsConfig.seed_eps.offset % 8 == 0


def test_ws_params_resolution():
    """which ws parameters the CUDA path accepts (the rest must raise, never silently run something else)"""
    import pytest
    from bootstrapper_b200.post.pipeline import resolve_ws_params
    p = resolve_ws_params({"bias": [-0.1, -0.2, -0.2], "seed_eps": 0.01, "fragments_in_xy": False})
    assert p["seed_eps"] == 0.01 and p["thresholds"] == [0.2, 0.35, 0.5]
    assert resolve_ws_params({"sigma": [1, 2, 2]})["sigma"] == [1, 2, 2]
    assert resolve_ws_params({"noise_eps": 0.001})["noise_seed"] == 0      # seeded stand-in for the reference's unseeded noise
    assert resolve_ws_params({"epsilon_agglomerate": 0.05})["epsilon_agglomerate"] == 0.05
    for bad in ({"epsilon_agglomerate": 0.05, "noise_eps": 0.01}, {"merge_function": "hist_quant_75"}):
        with pytest.raises(NotImplementedError):
            resolve_ws_params(bad)


def test_gaussian_weights_match_scipy():
    """the kernel handed to libbsnative is scipy's own (scipy.ndimage._filters._gaussian_kernel1d, order 0, truncate 4)"""
    import numpy as np
    from scipy.ndimage import _filters
    from bootstrapper_b200.native import gaussian_weights
    for sigma in (0.3, 0.7, 1.0, 2.0, 3.3, 8.0):
        r, w = gaussian_weights(sigma)
        assert r == int(4.0 * float(sigma) + 0.5)
        assert np.array_equal(w, _filters._gaussian_kernel1d(float(sigma), 0, r)[::-1])
    assert gaussian_weights(0)[0] == -1


def test_host_decoder_of_the_compact_form(built_lib):
    """bs_expand_compact is host code (no GPU): dense numbers + node-id table + LUT rows -> the uint64 arrays, background
    stays 0, any thread count gives the same arrays, a dense id beyond the table is refused"""
    import numpy as np
    import pytest
    import torch
    from bootstrapper_b200 import native
    rng = np.random.default_rng(3)
    n = 1000
    nodes = np.sort(rng.choice(10 ** 12, n, replace=False)).astype(np.int64) + 1
    luts = [nodes[rng.integers(0, n, n)] for _ in range(3)]
    dense = rng.integers(0, n + 1, (7, 33, 41)).astype(np.int32)
    want_f = np.where(dense > 0, nodes[np.maximum(dense, 1) - 1], 0)
    td, tn, tl = torch.from_numpy(dense), torch.from_numpy(nodes), [torch.from_numpy(l) for l in luts]
    for threads in (1, 5, 64):
        f, segs = native.expand_compact(td, tn, tl, threads=threads)
        assert np.array_equal(f.numpy(), want_f)
        for s_, l in zip(segs, luts):
            assert np.array_equal(s_.numpy(), np.where(dense > 0, l[np.maximum(dense, 1) - 1], 0))
    f0, s0 = native.expand_compact(td, tn, [], threads=2)          # fragments only
    assert np.array_equal(f0.numpy(), want_f) and s0 == []
    bad = td.clone()
    bad[0, 0, 0] = n + 1
    with pytest.raises(native.BsError):
        native.expand_compact(bad, tn, tl, threads=2)
