"""CPU test of the CTA-wide band solver of the refine kernel (csrc/pbcr_solver.cuh: short horizon partitions +
block cyclic reduction on the separators).  Its phases are plain host/device functions; tests/cpp/test_pbcr.cpp
runs them with a loop over the thread index standing in for the threads of a phase and compares the solution
with a dense Cholesky solve in extended precision on random banded SPD systems of 3 .. 512 time blocks
(all partition remainders, 0 .. 128 separators, 0 .. 8 reduction levels)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_pbcr_solver_matches_dense_cholesky():
    subprocess.run(["make", "-C", os.path.join(HERE, "cpp"), "test_pbcr"], check=True, capture_output=True)
    out = subprocess.run([os.path.join(HERE, "cpp", "test_pbcr")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "pbcr ok" in out.stdout
    assert float(out.stdout.split()[-1]) < 1e-12
