"""The drop-in boundary: the C-ABI library loads, exports every symbol the header declares, the
product never routes through the oracle or a CPU fallback, and fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "glasdi_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(glasdi_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from gplasdi_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/glasdi_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in gplasdi_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(declared)
    assert lib.glasdi_abi_version() == 1


def test_no_cuda_types_in_header():
    txt = open(os.path.join(ROOT, "include", "glasdi_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)          # signatures only, comments stripped
    for banned in ("#include <cuda", "torch", "at::", "cudaStream_t", "std::"):
        assert banned not in code, banned


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gplasdi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "reference_path" not in src, f


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    from gplasdi_b200 import _lib
    from gplasdi_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.glasdi_create(ctypes.byref(h), 0) == _lib.ERR_NO_DEVICE
    # decoder inference has no CPU fallback either
    from gplasdi_b200.latent_space import MultiLayerPerceptron
    mlp = MultiLayerPerceptron([5, 16, 64], "sigmoid", reshape_index=-1, reshape_shape=[64])
    with torch.no_grad(), pytest.raises(RuntimeError):
        mlp(torch.zeros(3, 5))
    # ... while the differentiable (training) evaluation is plain torch and works anywhere
    y = mlp(torch.zeros(3, 5, requires_grad=True))
    assert y.shape == (3, 64) and y.requires_grad


def test_state_dict_keys_match_reference_layout():
    from gplasdi_b200.latent_space import Autoencoder

    class Phys:
        qgrid_size = [33]

    ae = Autoencoder(Phys(), {"hidden_units": [12], "latent_dimension": 5})
    keys = sorted(ae.state_dict().keys())
    assert keys == ["decoder.fcs.0.bias", "decoder.fcs.0.weight", "decoder.fcs.1.bias", "decoder.fcs.1.weight",
                    "encoder.fcs.0.bias", "encoder.fcs.0.weight", "encoder.fcs.1.bias", "encoder.fcs.1.weight"]
    assert ae.decoder.layer_sizes == [5, 12, 33] and ae.n_z == 5


def test_registries_use_reference_type_keys():
    import gplasdi_b200
    assert set(gplasdi_b200.trainer_dict) == {"gplasdi"}
    assert set(gplasdi_b200.latent_dict) == {"ae"}
    assert set(gplasdi_b200.ld_dict) == {"sindy"}


def test_sindy_constructor_contract():
    from gplasdi_b200.latent_dynamics.sindy import SINDy
    ld = SINDy(5, 101, {"sindy": {"fd_type": "sbp24", "coef_norm_order": "fro"}})
    assert (ld.dim, ld.nt, ld.ncoefs, ld.fd_type) == (5, 101, 30, "sbp24")
    assert ld.export() == {"dim": 5, "ncoefs": 30, "fd_type": "sbp24", "coef_norm_order": "fro"}
    ld.load({"dim": 5, "ncoefs": 30})
    with pytest.raises(KeyError):
        SINDy(5, 101, {"sindy": {"fd_type": "weno"}})
    with pytest.raises(AssertionError):
        SINDy(5, 7, {"sindy": {"fd_type": "sbp48"}})      # shorter than the boundary closure
    # a non-uniform grid is accepted by simulate (glasdi_rollout_grid, tested on the GPU); the calls that need ONE step
    # (the fused greedy step, the exact propagator) still refuse it
    from gplasdi_b200.latent_dynamics.sindy import _uniform_dt, _is_uniform
    assert not _is_uniform(np.array([0.0, 0.1, 0.3])) and _is_uniform(np.linspace(0.0, 1.0, 11))
    with pytest.raises(ValueError):
        _uniform_dt(np.array([0.0, 0.1, 0.3]))


def _build_c_host(tmp_path):
    """gcc -std=c99 build of tests/c_host/abi_smoke.c against include/glasdi_b200.h and the in-tree library."""
    import shutil
    import subprocess
    from gplasdi_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc / CUDA runtime headers not available")
    _lib.load()
    libdir = os.path.dirname(_lib.lib_path())
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", os.path.join(root, "tests", "c_host", "abi_smoke.c"),
           "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
           "-L", libdir, "-lglasdi_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """The boundary is usable from a compiled C host: the header parses as C99 (-pedantic) and a C program that uses
    only the header + the CUDA runtime links against the library (it is RUN by tests/test_gpu_api.py on a B200)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", "-pedantic",
                           os.path.join(root, "include", "glasdi_b200.h")])
    exe = _build_c_host(tmp_path)
    assert os.path.exists(exe)


def test_trainer_export_load_round_trip_with_reference_keys(tmp_path):
    """export() carries every key of the reference trainer's dict (reference src/lasdi/gplasdi.py:276-283) -- optimiser
    state, timer dict and the four loss histories included -- and load() restores them, so a restart file written by
    either implementation loads in the other."""
    import torch
    from gplasdi_b200.latent_space import Autoencoder
    from gplasdi_b200.latent_dynamics.sindy import SINDy
    from gplasdi_b200.gplasdi import BayesianGLaSDI

    class Phys:
        qgrid_size = [33]; dt = 0.1; t_grid = np.linspace(0, 1, 11)

    class PS:
        train_space = np.zeros((2, 2)); test_space = np.zeros((3, 2))
        def n_train(self): return 2
        def n_test(self): return 3

    cfg = {"n_samples": 4, "lr": 1e-3, "n_iter": 2, "max_iter": 4, "max_greedy_iter": 1, "ld_weight": 0.1, "coef_weight": 1e-3,
           "path_checkpoint": str(tmp_path / "c"), "path_results": str(tmp_path / "r"), "device": "cpu"}

    def make():
        torch.manual_seed(3)
        ae = Autoencoder(Phys(), {"hidden_units": [12], "latent_dimension": 5})
        return BayesianGLaSDI(Phys(), ae, SINDy(5, 11, {"sindy": {"fd_type": "sbp12"}}), PS(), cfg), ae

    tr, ae = make()
    # two Adam steps on a plain torch loss populate the optimiser state (the GPU-only calibration is not needed for this)
    for _ in range(2):
        tr.optimizer.zero_grad()
        sum((p ** 2).sum() for p in ae.parameters()).backward()
        tr.optimizer.step()
    tr.X_train = torch.ones(2, 11, 33); tr.X_test = torch.ones(3, 11, 33)
    tr.best_coefs = np.arange(60.0).reshape(2, 30); tr.restart_iter = 2
    tr.training_loss, tr.ae_loss, tr.ld_loss, tr.coef_loss = [3.0, 2.0], [1.0, 0.9], [0.5, 0.4], [7.0, 6.0]
    d = tr.export()
    reference_keys = {"X_train", "X_test", "lr", "n_iter", "n_samples", "best_coefs", "max_iter", "ld_weight", "coef_weight",
                      "restart_iter", "timer", "optimizer", "training_loss", "ae_loss", "ld_loss", "coeff_loss"}
    assert reference_keys <= set(d)
    assert set(d["timer"]) == {"names", "calls", "times"}
    tr2, ae2 = make()
    tr2.load(d)
    assert tr2.restart_iter == 2 and np.array_equal(tr2.best_coefs, tr.best_coefs) and tr2.coef_loss == [7.0, 6.0]
    s1, s2 = tr.optimizer.state_dict()["state"], tr2.optimizer.state_dict()["state"]
    assert s1.keys() == s2.keys() and len(s1) > 0
    for k in s1:
        assert torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"]) and torch.equal(s1[k]["exp_avg_sq"], s2[k]["exp_avg_sq"])
        assert float(s1[k]["step"]) == float(s2[k]["step"]) == 2.0


def test_library_carries_the_hash_of_its_sources():
    """The loader refuses (or rebuilds) a library built from other sources: the sha1 of csrc/ + the header is compiled into
    the library and written beside it (gplasdi_b200/build.py::source_hash)."""
    from gplasdi_b200 import _lib, build
    lib = _lib.load()
    assert lib.glasdi_source_hash().decode() == build.source_hash() == build.built_hash()
    assert not build.needs_build()
