"""CPU tests of the solution-file writer/readers (dumpSolutions format,
inter_agent_cons.cc:413-455, and the status parse of scripts/analysis_result.py:53-101)."""
import numpy as np

from csdotrajectoryplanning_b200.output import (SolutionStatistics, format_solutions, dump_solutions, load_solutions,
                                                read_solution_status)


def _traj():
    tr = np.zeros((2, 6, 3))
    tr[0, 0] = [1.0, 1.70649, 2.4125]
    tr[0, 1] = [2.0, 2.0005, -0.00049]
    tr[0, 2] = [0.5, 0.5, 0.5]
    tr[0, 3] = [0.0, 0.321750554, -0.321750554]
    tr[0, 4] = [0.8, 0.8, 0.0]
    tr[0, 5] = [0.07, -0.07, 0.0]
    tr[1, 0] = [10, 10, 10]
    return tr


def test_format_matches_dump_solutions_layout():
    txt = format_solutions(_traj(), SolutionStatistics(rt_search=0.25, solver_status=-2, search_status=1))
    lines = txt.splitlines()
    assert lines[:12] == ["statistics:", "  cost: -1.000", "  makespan: -1.000", "  flowtime: -1.000", "  runtime: -1.000",
                          "  runtime_search: 0.250", "  runtime_preprocess: -1.000", "  runtime_optimization: -1.000",
                          "  runtime_decentralized_optimization: -1.000", "  search_status: 1", "  solver_status: -2",
                          "schedule:"]
    assert lines[12] == "  agent0:"
    # first step: x, y, yaw, steer (scaled by 180/3.14), t, v, omega
    assert lines[13:20] == ["    - x: 1.000", "      y: 2.000", "      yaw: 0.500", "      steer: 0.000", "      t: 0",
                            "      v: 0.800", "      omega: 4.013"]
    # second step: std::fixed rounding of the binary value, steer in pseudo-degrees
    assert lines[20] == "    - x: 1.706" and lines[23] == "      steer: 18.444"
    # the last step carries no v / omega (inter_agent_cons.cc:449)
    last = lines[27:32]
    assert last == ["    - x: 2.413", "      y: -0.000", "      yaw: 0.500", "      steer: -18.444", "      t: 2"]
    assert lines[32] == "  agent1:"
    assert txt.endswith("      t: 2\n")


def test_round_trip_and_status_rule(tmp_path):
    p = str(tmp_path / "sol.yaml")
    tr = _traj()
    dump_solutions(p, tr, SolutionStatistics(solver_status=2, rt_preprocess=0.5))
    back = load_solutions(p)
    assert back.shape == tr.shape
    assert np.abs(back[:, :3] - tr[:, :3]).max() <= 5e-4 + 1e-12
    assert np.abs(back[:, 3] - tr[:, 3]).max() <= 5e-4 / (180 / 3.14) + 1e-12
    assert np.all(back[:, 4:, -1] == 0)
    st, ok = read_solution_status(p)
    assert ok and st.solver_status == 2 and st.rt_preprocess == 0.5 and st.search_status == 2
    for code, want in ((1, True), (-2, True), (2, True), (3, False), (-3, False), (-7, False)):
        dump_solutions(p, tr, SolutionStatistics(solver_status=code))
        assert read_solution_status(p)[1] == want


def test_cpp_dump_solutions_writes_the_same_file(tmp_path):
    """include/csdo/solution_io.h (C++) and output.format_solutions (Python) produce identical text."""
    import os, subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.run(["make", "-C", os.path.join(here, "cpp"), "test_solution_io"], check=True, capture_output=True)
    rng = np.random.default_rng(3)
    tr = rng.normal(0, 20, (3, 6, 7))
    tr[:, 3] *= 0.01; tr[:, 5] *= 0.003
    tr[0, 1, 2] = -0.0004          # prints as -0.000 in both
    stat = SolutionStatistics(solver_status=-2, search_status=1, rt_preprocess=0.125)
    lines = [f"3 7 {stat.solver_status} {stat.search_status} {stat.rt_preprocess!r}"]
    for a in range(3):
        for t in range(7):
            lines.append(" ".join(repr(float(tr[a, k, t])) for k in range(6)))
    path = str(tmp_path / "cpp.yaml")
    subprocess.run([os.path.join(here, "cpp", "test_solution_io"), path], input="\n".join(lines) + "\n", text=True,
                   check=True)
    assert open(path).read() == format_solutions(tr, stat)


def test_dump_corridors_cpp_and_python_agree(tmp_path):
    """dumpCorridors (sqp/utils.cc:62-89; restated, the file itself needs Eigen/OSQP headers so it is not part of
    oracle/_ref): C++ header and Python writer give the same text, with a hand-checked first row."""
    import os, subprocess
    from csdotrajectoryplanning_b200 import default_params
    from csdotrajectoryplanning_b200.output import format_corridors
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.run(["make", "-C", os.path.join(here, "cpp"), "test_solution_io"], check=True, capture_output=True)
    p = default_params()
    rng = np.random.default_rng(5)
    na, nt = 2, 5
    guess = np.stack([rng.uniform(0, 100, (na, nt)), rng.uniform(0, 100, (na, nt)), rng.uniform(-3.2, 3.2, (na, nt))], axis=1)
    guess[0, :, 0] = (10.0, 20.0, 0.0)                    # front disc at (11.25, 20), rear at (9.75, 20)
    corr = guess[:, [0, 0, 1, 1, 0, 0, 1, 1], :] + rng.uniform(-10, 10, (na, 8, nt))
    corr[0, :, 0] = (8.5, 14.0, 17.123456789, 23.0, 1e-7, 12.0, 17.0, 1234567.0)
    lines = [f"{na} {nt} {p.f2x!r} {p.r2x!r}"]
    for a in range(na):
        for t in range(nt):
            lines.append(" ".join(repr(float(v)) for v in list(guess[a, :, t]) + list(corr[a, :, t])))
    path = str(tmp_path / "corridors.yaml")
    subprocess.run([os.path.join(here, "cpp", "test_solution_io"), path, "corridors"], input="\n".join(lines) + "\n",
                   text=True, check=True)
    text = format_corridors(corr, guess, p.f2x, p.r2x)
    assert open(path).read() == text
    rows = text.split("\n")
    assert rows[0] == "agent0:"
    assert rows[1] == "  - [11.25, 20, 8.5, 14, 17.1235, 23]"
    assert rows[2] == "  - [9.75, 20, 1e-07, 12, 17, 1.23457e+06]"
