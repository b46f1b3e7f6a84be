"""Static properties of the shipped sm_100a code that the measured numbers depend on
(profiles/README.md, rounds 1e/1f).  CPU-only: reads the SASS of the in-tree library with
cuobjdump; nothing is executed."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gerris-fft-particles_b200", "lib", "libgfsb200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB),
                                reason="needs cuobjdump and the built library")


@pytest.fixture(scope="module")
def sass():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            funcs[name].append(line)
    return funcs


def _one(funcs, pattern):
    hits = [k for k in funcs if re.search(pattern, k)]
    assert len(hits) == 1, (pattern, hits)
    return "\n".join(funcs[hits[0]])


def test_library_is_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


# the C2 / C3 production instances: 3D, lattice or not, drag+lift+buoyancy, 2 stages, 7 CTAs x 4 warps
@pytest.mark.parametrize("lattice", [1, 0])
def test_warp_pipelined_step_kernel(sass, lattice):
    # (last flag, round 2: the stream policy -- ordinary loads / stores on the big uniform tree,
    # evict-first + streaming stores on the adaptive ones)
    code = _one(sass, r"step_kernel_wpipeILi3ELb%dELj801ELi2ELi7ELi4ELb0ELb0ELb%dE" % (lattice, 1 - lattice))
    # the particle stream comes in through the TMA engine onto mbarriers
    assert code.count("UBLKCP") == 3 * 8            # prologue (2 stages) + refill, 8 columns each
    assert "SYNCS.PHASECHK.TRANS64.TRYWAIT" in code and "SYNCS.ARRIVE.TRANS64" in code
    # ... one UBLKCP per copy: no lane-serialising loop around it (round 1f)
    assert "BRA.U.ANY" not in code
    assert code.count("R2UR") <= 4
    assert code.count("ELECT") <= 3
    # no CTA-wide barrier in the loop, nothing in local memory (72 registers, no spills)
    assert "BAR.SYNC" not in code
    assert not re.search(r"\b(LDL|STL)\b", code)
    # velocity, mass, volume and the position are fetched from the staged tile with LDS.64
    assert len(re.findall(r"\bLDS\.64\b", code)) >= 11


@pytest.mark.parametrize("lattice", [1, 0])
def test_fused_step_deposit_kernel(sass, lattice):
    """the two-way flavour of the same kernel (round 2): same TMA pipeline, no spills, and the
    deposits leave as fp64 reductions without a return value (a remote owner costs no round trip)"""
    code = _one(sass, r"step_kernel_wpipeILi3ELb%dELj801ELi2ELi7ELi4ELb0ELb1ELb1E" % lattice)
    assert code.count("UBLKCP") == 3 * 8
    assert "BAR.SYNC" not in code
    assert not re.search(r"\b(LDL|STL)\b", code)
    # void fraction + three force components, once at GPU scope (one rank) and once at system
    # scope (the slice also receives reductions from peers over NVLink)
    reds = re.findall(r"\bREDG\.E\.ADD\.F64\.RN\.STRONG\.(GPU|SYS)", code)
    assert sorted(reds) == ["GPU"] * 4 + ["SYS"] * 4, reds
    assert not re.search(r"\bATOMG?\.E\.ADD\.F64", code)


def test_force_recording_step_kernel(sass):
    """the flavour the GModule uses (GfsParticulate.force is part of the object): three more
    column stores on the same pipeline"""
    code = _one(sass, r"step_kernel_wpipeILi3ELb1ELj801ELi2ELi7ELi4ELb1ELb0ELb1E")
    assert code.count("UBLKCP") == 3 * 8
    assert not re.search(r"\b(LDL|STL)\b", code)
    # six state columns (ordinary stores since round 2: measured faster in the kernels without the
    # deposit tail) + three force columns (streaming)
    assert len(re.findall(r"\bSTG\.E(\.EF)?\.64\b", code)) >= 9
    assert len(re.findall(r"\bSTG\.E\.EF\.64\b", code)) >= 3


def test_cell_pass_shared_loads(sass):
    code = _one(sass, r"lattice_cell_pass_kernelILi2206526ELb1E")
    # round 2: the 512 leaves of the brick arrive as three bulk copies, only the halo as gathers
    assert code.count("UBLKCP") == 3
    # 2 x (24 region cells + 6 "-" neighbours) 64-bit shared loads per thread (round 1e)
    # (+ 6: two cells x three components moved from the bulk staging array into the region; + 12
    # STATIC ones: the gradient block exists twice, for bricks off the hull without the case
    # selection and for the others -- a thread executes one of them)
    assert len(re.findall(r"\bLDS\.64\b", code)) <= 78
    assert len(re.findall(r"\bLDGSTS\b", code)) <= 6
    assert "LDGSTS" in code
    # one 4-byte value parked across the compute phase for the hull part: nothing in the loop
    assert len(re.findall(r"\b(LDL|STL)\b", code)) <= 4
