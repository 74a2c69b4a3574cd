"""TEST INFRASTRUCTURE ONLY.  Link-level drop-in check.

Builds the reference's UNMODIFIED Cython binding (/root/reference/utils/csegment/c_segment.pyx, copied to a
scratch directory: the reference tree is read-only and its sources never enter this repo) with `segment.cc`
dropped from `sources` and `libmergenet_b200.so` linked instead, exactly the setup.py edit INTEGRATION.md
shows.  The built extension module goes to oracle/_ref/cython_dropin/ (git-ignored, travels to the GPU box
like the other built .so files), where tests/test_cython_dropin.py imports it.

    python oracle/build_cython_dropin.py
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference/utils/csegment"
OUT = os.path.join(HERE, "_ref", "cython_dropin")

SETUP = '''
from setuptools import setup
from setuptools.extension import Extension
from Cython.Build import cythonize
import numpy
setup(ext_modules=cythonize([Extension(
    "c_segment", sources=["c_segment.pyx"], language="c++",
    include_dirs=[numpy.get_include()],
    libraries=["mergenet_b200"], library_dirs=[{libdir!r}],
    extra_link_args=["-Wl,-rpath,$ORIGIN/../../../mergenet_b200", "-Wl,-rpath,{libdir}"],
    extra_compile_args=["-std=c++11"])], language_level=2))
'''


def build(verbose=False):
    if not os.path.exists(os.path.join(REF, "c_segment.pyx")):
        return None
    libdir = os.path.join(REPO, "mergenet_b200")
    if not os.path.exists(os.path.join(libdir, "libmergenet_b200.so")):
        raise RuntimeError("build libmergenet_b200.so first (__graft_entry__.build())")
    tmp = tempfile.mkdtemp(prefix="mn_cython_")
    try:
        shutil.copy(os.path.join(REF, "c_segment.pyx"), tmp)  # unmodified
        with open(os.path.join(tmp, "setup.py"), "w") as f:
            f.write(SETUP.format(libdir=libdir))
        env = dict(os.environ)
        # Cython 3 declares `cdef extern` prototypes extern "C++" in C++ mode while the symbol
        # is extern "C" (segment.cc:742 and include/mergenet_b200.h alike); link with g++ so libstdc++ comes along
        env["CPPFLAGS"] = env.get("CPPFLAGS", "") + " -DCYTHON_EXTERN_C='extern \"C\"'"
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        env.update(CC=cc, CXX=cxx, LDSHARED=cxx + " -shared", LDCXXSHARED=cxx + " -shared")
        out = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, env=env,
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or out.returncode:
            print(out.stdout)
        if out.returncode:
            raise RuntimeError("cython drop-in build failed")
        mods = glob.glob(os.path.join(tmp, "c_segment*.so"))
        assert mods, "no extension module built"
        os.makedirs(OUT, exist_ok=True)
        for old in glob.glob(os.path.join(OUT, "c_segment*.so")):
            os.remove(old)
        dst = os.path.join(OUT, os.path.basename(mods[0]))
        shutil.copy(mods[0], dst)
        return dst
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(build(verbose=True))
