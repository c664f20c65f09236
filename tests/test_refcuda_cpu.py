"""CPU-side checks of the reference-device-code harness (oracle/ref_cuda): it builds from the reference checkout
when that is present, the libraries load and export the harness ABI, and the oracle's pin-test switch is what
it says.  No GPU work here."""
import os

import numpy as np
import pytest

from _common import bits, scene
from oracle import oracle_py as O
from oracle import refcuda_py as R


@pytest.mark.parametrize("flavour", ["ieee", "fast"])
def test_harness_library_loads_and_exports(flavour):
    if not R.available(flavour):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref not built and no reference checkout on this box")
        import subprocess

        subprocess.run(["make", "-C", os.path.join(os.path.dirname(R.REF_DIR), "ref_cuda"), "-s"], check=True)
    L = R.lib(flavour)
    for s in R.SYMBOLS:
        assert hasattr(L, s), s


def test_nothing_of_the_reference_is_copied_into_the_repo():
    here = os.path.join(os.path.dirname(R.REF_DIR), "ref_cuda")
    assert sorted(os.listdir(here)) == ["Makefile", "ref_harness.cu"]
    src = open(os.path.join(here, "ref_harness.cu")).read()
    assert '#include "bvh.cu"' in src  # compiled from where it lies, via -I


def test_sequential_switch_changes_only_the_association():
    sc = scene("tiny")
    rng = np.random.default_rng(3)
    o = rng.uniform(-9, 9, size=(4096, 3)).astype(np.float32)
    o[:, 2] = rng.uniform(0.5, 6, size=4096)
    d = rng.normal(size=(4096, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    a = sc.oracle.intersect_rays_raw(o, d, 0.0, 60.0)
    O.set_sqnorm_sequential(True)
    try:
        b = sc.oracle.intersect_rays_raw(o, d, 0.0, 60.0)
    finally:
        O.set_sqnorm_sequential(False)
    c = sc.oracle.intersect_rays_raw(o, d, 0.0, 60.0)
    assert np.array_equal(a["leaf"], c["leaf"]) and np.array_equal(bits(a["dsq"]), bits(c["dsq"]))
    assert np.array_equal(a["leaf"] >= 0, b["leaf"] >= 0)
    db = np.abs(bits(a["dsq"]).astype(np.int64) - bits(b["dsq"]).astype(np.int64))
    assert db.max() <= 1 and (a["leaf"] != b["leaf"]).mean() < 1e-2
    # raw == normalising entry point when the directions are already unit length up to re-normalisation rounding
    e = sc.oracle.intersect_rays(o, d, 0.0, 60.0)
    assert (e["leaf"] == a["leaf"]).mean() > 0.99
