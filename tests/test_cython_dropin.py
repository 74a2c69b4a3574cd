"""Link-level drop-in: the reference's UNMODIFIED Cython binding
(/root/reference/utils/csegment/c_segment.pyx:16-25,30-86), built with `segment.cc` dropped from `sources`
and libmergenet_b200.so linked instead (oracle/build_cython_dropin.py; setup.py:11-16 edited as INTEGRATION.md
shows).  CPU: the module builds, resolves `c_run_segmentation` from our library, keeps the wrapper's argument
checks, and fails loudly without a device.  GPU: calling through the reference's own wrapper gives the
oracle's result."""
import glob
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "cython_dropin")


def _module_path(build_if_possible):
    mods = glob.glob(os.path.join(DROPIN, "c_segment*.so"))
    if build_if_possible and os.path.exists("/root/reference/utils/csegment/c_segment.pyx"):
        lib = os.path.join(ROOT, "mergenet_b200", "libmergenet_b200.so")
        if not mods or os.path.getmtime(mods[0]) < os.path.getmtime(lib):
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import build_cython_dropin
            build_cython_dropin.build()
            mods = glob.glob(os.path.join(DROPIN, "c_segment*.so"))
    return mods[0] if mods else None


def _load(path):
    spec = importlib.util.spec_from_file_location("c_segment", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_unmodified_cython_binding_links_against_the_library(lib_mod):
    path = _module_path(True)
    if path is None:
        pytest.skip("reference tree absent and no prebuilt binding")
    syms = subprocess.run(["nm", "-D", path], stdout=subprocess.PIPE, text=True).stdout
    assert " U c_run_segmentation" in syms          # unmangled, undefined here: comes from the library
    needed = subprocess.run(["readelf", "-d", path], stdout=subprocess.PIPE, text=True).stdout
    assert "libmergenet_b200.so" in needed
    cseg = _load(path)
    name, cp, sp, C, offs = cases.small_cases()[0]
    with pytest.raises(TypeError):                    # pyx:30-31 "not None"
        cseg.run_segmentation(None, sp, C, offs, 0.0, 1.0, 0.03)
    with pytest.raises(ValueError):                   # buffer dtype check of the typed argument
        cseg.run_segmentation(cp.astype(np.float64), sp, C, offs, 0.0, 1.0, 0.03)
    if lib_mod.lib().mn_device_count() == 0:
        # no device: the void symbol reports on stderr and leaves the outputs empty; no silent CPU path
        mask, ocls = cseg.run_segmentation(cp.copy(), sp.copy(), C, offs, 0.0, 1.0, 0.03)
        assert not mask.any() and ocls == []
        assert lib_mod.lib().mn_last_error() == 7


@pytest.mark.gpu
def test_reference_wrapper_over_the_cuda_library_matches_oracle(oracle_mod, lib_mod):
    path = _module_path(False)
    if path is None:
        pytest.skip("oracle/_ref/cython_dropin not built (python oracle/build_cython_dropin.py in the build container)")
    cseg = _load(path)
    for name, cp, sp, C, offs in cases.small_cases()[:4]:
        for opts in (cases.RECIPE_OPTS, (0.5, 1.0, 0.0)):
            sp1 = sp.copy()
            mask, ocls = cseg.run_segmentation(cp.copy(), sp1, C, offs, *opts)
            assert lib_mod.lib().mn_last_error() == 0
            m0, c0, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
            assert mask.dtype == np.int32 and cases.same_result(oracle_mod, (m0, c0), (mask, ocls)), (name, opts)
