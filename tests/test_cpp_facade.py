"""The C++ facade (vgan_b200/csrc/host/vgan_haplocart.hpp) built into a small host program shaped like the patched
src/HaploCart.cpp; its output must agree with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import vgan_b200 as vb
from oracle import oracle as orc

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "facade_demo.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "facade_demo")


def _build():
    lib = os.path.join(ROOT, "vgan_b200", "lib")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(
            os.path.join(ROOT, "vgan_b200", "csrc", "host", "vgan_haplocart.hpp"))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", EXE, SRC, "-L" + lib, "-lvgan_host", "-lvgan_gpu",
                               "-Wl,-rpath," + lib])
    return EXE


def test_facade_compiles_and_flattens_alignment_objects():
    out = subprocess.check_output([_build(), "64", "--no-gpu"]).decode()
    assert "flattened 64 reads" in out


@pytest.mark.gpu
@pytest.mark.parametrize("route", [[], ["--device-reconstruction"], ["--streamed"], ["--samples"], ["--in-flight"]])
def test_facade_end_to_end_matches_oracle(route):
    out = subprocess.check_output([_build(), "200"] + route).decode().splitlines()
    S = vb.SynthHaplocart(seed=101, n_paths=300, n_rows=1210, n_nodes=1201)
    alns = S.reads(200, seed=202)
    T, G, R = H.oracle_side(S, alns.arrays())
    final, argmax, _ = orc.hc_run(T, G, R)
    call = out[0].split()
    assert int(call[1]) == argmax and call[2] == S.path_names[argmax]
    ll = np.array([float(l.split()[2]) for l in out if l.startswith("ll ")])
    assert H.rel_err(ll, final.astype(np.float64)) < 1e-11
    idx, conf, total = orc.hc_posterior(final, S.path_names, S.parents_txt, S.children_txt, argmax)
    got = [float(l.split()[2]) for l in out if l.startswith("clade ")]
    assert np.allclose(got, conf, rtol=1e-6, atol=1e-300)
    assert float(call[4]) == pytest.approx(float(total), rel=1e-12)
