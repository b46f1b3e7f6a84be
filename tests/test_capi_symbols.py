"""CPU tests of the drop-in boundary: every function declared in include/*.h is
exported by the built libraries, and the compute entry points fail loudly
(no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

import helpers
from helpers import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gfsb200_\w+)\s*\(", text)) - {"gfsb200_refine_func"})


def test_libgfsb200_exports_every_declared_symbol():
    lib = capi.lib()
    names = declared("gfsb200.h")
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gfsb200.h but not exported"


@pytest.mark.parametrize("dim", [2, 3])
def test_bridge_exports_every_declared_symbol(dim):
    b = capi.bridge(dim)
    for n in declared("gfsb200_ftt.h"):
        assert hasattr(b, n), n


def test_no_cpu_fallback():
    """without a CUDA device ctx_create must fail with GFSB200_ERR_CUDA"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = capi.lib().gfsb200_ctx_create(0, ctypes.byref(h))
    assert rc == -4 and not h.value
    assert b"no CPU path" in capi.lib().gfsb200_last_error()
    with pytest.raises(capi.GfsB200Error):
        capi.Context(0)


def test_product_does_not_touch_the_oracle():
    """nothing under gerris-fft-particles_b200/ or include/ refers to oracle/"""
    bad = []
    for base in ("gerris-fft-particles_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".c", ".cu", ".cuh", ".h", ".py")):
                    src = open(os.path.join(d, f), errors="ignore").read()
                    if re.search(r"oracle[/.]|libgfsoracle|ora_[a-z_]+\s*\(", src):
                        bad.append(os.path.join(d, f))
    # the Makefile may point GTS_INC at the shim to compile-check the bridge; sources may not
    assert not bad, bad


def test_kernel_fit_recognises_the_closed_forms():
    """gfsb200_kernel_fit (host only): a user kernel GfsFunction is recognised by
    probing; anything that is not one of the device's closed forms is refused so
    that the module keeps such a list on the reference's CPU event."""
    import math
    import pytest
    from helpers import capi
    k = capi.kernel_fit(lambda x, y, z: 0.25, 3)
    assert (k.kind, k.a) == (capi.KERNEL_CONSTANT, 0.25)
    k = capi.kernel_fit(lambda x, y, z: 3.0 * math.exp(-(x * x + y * y + z * z) / 2.0), 3)
    assert k.kind == capi.KERNEL_GAUSSIAN and k.a == 3.0 and abs(k.b - 0.5) < 1e-14
    k = capi.kernel_fit(lambda x, y, z: math.exp(-0.02 * (x * x + y * y)), 2)
    assert k.kind == capi.KERNEL_GAUSSIAN and abs(k.b - 0.02) < 1e-14
    for p in (1, 2, 4):
        k = capi.kernel_fit(lambda x, y, z: 2.0 * max(0.0, 1.0 - (x * x + y * y + z * z) / 9.0) ** p, 3)
        assert (k.kind, k.p, k.a) == (capi.KERNEL_COMPACT, p, 2.0) and abs(k.b - 1.0 / 9.0) < 1e-12
    for bad in (lambda x, y, z: math.exp(-abs(x) - abs(y) - abs(z)),        # not radial-quadratic
                lambda x, y, z: math.exp(-(x * x + 2 * y * y + z * z)),      # anisotropic
                lambda x, y, z: 1.0 / (1.0 + x * x + y * y + z * z),
                lambda x, y, z: x * x + y * y,                               # zero at the centre
                lambda x, y, z: float("nan")):
        with pytest.raises(capi.GfsB200Error):
            capi.kernel_fit(bad, 3)


def test_dropin_module_type_checks_against_reference_headers():
    """host/particulates_b200.c (the libparticulates{2D,3D}.so replacement) is type-checked
    with gcc -fsyntax-only against the reference's own src/*.h and
    modules/particulatecommon.h, GLib/GTS being stood in for by declaration-only headers
    (host/check/).  Needs the reference tree; skipped on the GPU box."""
    import os
    import subprocess
    import pytest
    if not os.path.exists("/root/reference/src/ftt.h"):
        pytest.skip("reference tree not present")
    env = dict(os.environ)
    env.pop("CC", None)
    pkg_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gerris-fft-particles_b200")
    r = subprocess.run(["make", "check-host"], cwd=pkg_dir, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "warning" not in r.stdout, r.stdout[-3000:]
