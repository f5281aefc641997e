"""The C-ABI library: it loads, exports every symbol include/smcmc_b200.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "smcmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smcmc_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    import smcmc_b200
    from smcmc_b200 import binding
    if not os.path.exists(smcmc_b200.library_path()):
        smcmc_b200.build_library()
    lib = ctypes.CDLL(smcmc_b200.library_path())
    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(binding.EXPORTED_SYMBOLS) == names
    assert lib.smcmc_abi_version() == 2


def test_struct_layouts_match_reference_event():
    import smcmc_b200
    assert smcmc_b200.EVENT_DTYPE.itemsize == 48          # sizeof(Simulated::Event)
    assert smcmc_b200.EVENT_DTYPE.fields["Separation"][1] == 16
    assert smcmc_b200.EVENT_DTYPE.fields["TrueMass"][1] == 32


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import smcmc_b200
    with pytest.raises(smcmc_b200.SmcmcError) as err:
        smcmc_b200.Engine(smcmc_b200.LLH_FAKE, 9, 8)
    assert err.value.status == -5
    assert "no CPU fallback" in str(err.value)


def test_bad_arguments_are_rejected():
    import smcmc_b200
    from smcmc_b200 import binding
    lib = smcmc_b200.load_library()
    h = ctypes.c_void_p()
    cfg = binding._Config(ctypes.sizeof(binding._Config), 0, 0, 4, 0, 0, 1)    # dim 0
    assert lib.smcmc_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    cfg = binding._Config(4, 0, 3, 4, 0, 0, 1)                                 # wrong struct size
    assert lib.smcmc_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    cfg = binding._Config(ctypes.sizeof(binding._Config), 0, 5, 4, 0, 4, 1)    # Fake needs 9 dims
    assert lib.smcmc_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"9 parameters" in lib.smcmc_last_error(None)
