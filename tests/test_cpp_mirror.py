"""Runs the C++ mirror of the reference's Rust API (include/quickdna.hpp) against the reference's own
Rust unit-test vectors, through the C ABI, on the GPU.  The binary is built by __graft_entry__.build()."""

import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_rust_api_mirror")


def test_mirror_binary_is_built():
    import __graft_entry__ as g

    g.build()
    assert os.path.exists(BIN)


def test_lazy_adaptor_views_on_host():
    """NucleotideView / CodonsView / to_fn (src/iter.rs, src/trans_table.rs:165-174): host-side views, no kernel."""
    import __graft_entry__ as g

    g.build()
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "test_iter_views")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.mark.gpu
def test_rust_api_mirror_on_gpu():
    import __graft_entry__ as g

    g.build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.mark.gpu
def test_lazy_adaptor_bulk_forms_on_gpu():
    """a11 / a24: `frame.map(table.to_fn()).collect()` of every reading frame, to_fn over a slice of codons and the
    reverse-complement strand as ONE launch each, equal to the per-element host views."""
    import __graft_entry__ as g

    g.build()
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "test_iter_gpu")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
