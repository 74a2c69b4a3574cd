"""libm parity: the fp64 recipe behind the CUDA edge pass must reproduce the
host libm bit for bit on the whole clipped domain [2^-23, 1-2^-23] = float bits 0x34000000..0x3f7ffffe.
CPU: the oracle's restatement of the recipe vs libm (exhaustive).  GPU: the device functions vs host
tables produced by the oracle library (exhaustive, chunked)."""
import ctypes

import numpy as np
import pytest

LO_BITS = 0x34000000
HI_BITS = 0x3F7FFFFE


def test_logf_recipe_equals_libm_exhaustive(oracle_mod):
    bad = oracle_mod.oracle_lib().mno_logf_recipe_mismatches(LO_BITS, HI_BITS, 1)
    assert bad == 0


def test_expf_recipe_equals_libm_sampled(oracle_mod):
    """The same_different_bias path (segment.cc:189-191) needs expf; the recipe must be the host's expf."""
    L = oracle_mod.oracle_lib()
    L.mno_expf_recipe_mismatches.restype = ctypes.c_longlong
    L.mno_expf_recipe_mismatches.argtypes = [ctypes.c_uint32]
    assert L.mno_expf_recipe_mismatches(7) == 0


def test_log1m_64bin_table_exhaustive(tmp_path):
    """The 64-bin log(1 - s) evaluation of mn_edge_warp_kernel (table from tools/gen_log1m_table.py, degree-6
    polynomial, ambiguity test), emulated on the host with fma(): its float rounding must equal the host
    libm's (float)log(1.0 - (double)s) on every input of the clipped domain outside the fallback set."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "check_log1m64")
    flags = ["-O2"]
    try:
        if " fma " in open("/proc/cpuinfo").read():
            flags.append("-mfma")   # hardware fma: 3 s instead of a software fma() per step
    except OSError:
        pass
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc] + flags + [os.path.join(here, "check_log1m64.c"), "-lm", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "mismatches outside the fallback set 0" in out.stdout and "inputs 192937983" in out.stdout, out.stdout


def _device_vs_host(oracle_mod, lib_mod, which, host_fn, stride_chunks=1, bias=0.0):
    L = lib_mod.lib()
    F = ctypes.POINTER(ctypes.c_float)
    chunk = 1 << 24
    bad = 0
    first_bad = []
    idx = 0
    for start in range(LO_BITS, HI_BITS + 1, chunk):
        idx += 1
        if stride_chunks > 1 and idx % stride_chunks:
            continue
        n = min(chunk, HI_BITS + 1 - start)
        dev = np.empty(n, np.float32)
        host = np.empty(n, np.float32)
        st = L.mn_debug_libm(which, start, n, float(bias), dev.ctypes.data_as(F))
        assert st == 0
        if which == 2:
            host_fn(start, n, float(bias), host.ctypes.data_as(F))
        else:
            host_fn(start, n, host.ctypes.data_as(F))
        neq = dev.view(np.uint32) != host.view(np.uint32)
        c = int(neq.sum())
        if c:
            bad += c
            i = int(np.flatnonzero(neq)[0])
            first_bad.append((hex(start + i), float(dev[i]), float(host[i])))
    return bad, first_bad


@pytest.mark.gpu
def test_device_logf_equals_host_exhaustive(oracle_mod, lib_mod):
    bad, first = _device_vs_host(oracle_mod, lib_mod, 0, oracle_mod.oracle_lib().mno_host_logf_table)
    assert bad == 0, first[:5]


@pytest.mark.gpu
def test_device_log1m_equals_host_exhaustive(oracle_mod, lib_mod):
    bad, first = _device_vs_host(oracle_mod, lib_mod, 1, oracle_mod.oracle_lib().mno_host_log1m_table)
    assert bad == 0, first[:5]


@pytest.mark.gpu
def test_device_warp_pipeline_logs_equal_host_exhaustive(oracle_mod, lib_mod):
    """The (k, i)-table logf and the split log1m of mn_edge_warp_kernel, whole clipped domain."""
    L = oracle_mod.oracle_lib()
    bad, first = _device_vs_host(oracle_mod, lib_mod, 5, L.mno_host_logf_table)
    assert bad == 0, first[:5]
    bad, first = _device_vs_host(oracle_mod, lib_mod, 6, L.mno_host_log1m_table)
    assert bad == 0, first[:5]


@pytest.mark.gpu
def test_device_unfused_recipes_equal_host_sampled(oracle_mod, lib_mod):
    """The unfused logf recipe and the plain fp64 log (used by the same_different_bias path)."""
    L = oracle_mod.oracle_lib()
    bad, first = _device_vs_host(oracle_mod, lib_mod, 3, L.mno_host_logf_table, stride_chunks=3)
    assert bad == 0, first[:5]
    bad, first = _device_vs_host(oracle_mod, lib_mod, 4, L.mno_host_log1m_table, stride_chunks=3)
    assert bad == 0, first[:5]


@pytest.mark.gpu
def test_device_bias_transform_equals_host_sampled(oracle_mod, lib_mod):
    # same_different_bias != 0 (segment.cc:183-195): every 4th 16M-chunk of the domain, two biases
    for bias in (0.5, -1.25):
        bad, first = _device_vs_host(oracle_mod, lib_mod, 2, oracle_mod.oracle_lib().mno_host_bias_table,
                                     stride_chunks=4, bias=bias)
        assert bad == 0, (bias, first[:5])
