"""CPU test: the documented drop-in compiles INSIDE the reference tree (SURVEY section 8 row b).

The translation unit is generated at test time from the reference's own csdo.cc: its pre-process + DSQP
block (csdo.cc:111-168, everything after the PBS search) is wrapped in a function and compiled against the
reference's REAL headers (common/motion_planning.h, sqp/common.h, sqp/corridor.h, sqp/inter_agent_cons.h,
hybrid_a_star/*.h) with exactly the edit INTEGRATION.md documents:  `#include "sqp/dsqp_solver.h"` ->
`#define CSDO_WITH_REFERENCE_TYPES` + `#include "csdo/dsqp_solver.h"`.  Built twice with -fsyntax-only:
(a) the block unchanged (CPU plane functions of the reference, GPU SolverDSQP), (b) the two plane calls
qualified with csdo_b200:: (GPU plane kernels).  Needs /root/reference (skipped on the GPU box); Boost is
absent offline, so the two Boost headers motion_planning.h includes come from oracle/ref_build/stub.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

PRELUDE = r'''
#define CSDO_WITH_REFERENCE_TYPES
#include <array>
#include <cassert>
#include <iostream>
#include <string>
#include <unordered_set>
#include <vector>
#include "hybrid_a_star/timer.h"
#include "hybrid_a_star/planresult.h"
#include "common/motion_planning.h"
using Path = libMultiRobotPlanning::PlanResult<State, Action, double>;   // hybrid_a_star/types.h:14 (that header pulls OMPL)
#include "sqp/common.h"
#include "sqp/inter_agent_cons.h"
#include "sqp/corridor.h"
#include "csdo/dsqp_solver.h"        // replaces  #include "sqp/dsqp_solver.h"
using namespace libMultiRobotPlanning;
using std::cout; using std::endl; using std::string; using std::vector;
// what the block uses from the parts of csdo.cc that stay as they are (outside the path)
namespace libMultiRobotPlanning {
void readQpSolverConfig(std::string fname_config, QpParm& param);                       // sqp/utils.h:25
void dumpCorridors(std::string file_name, const std::vector<std::vector<Corridor>>& corridors,
                   const std::vector<std::vector<OptimizeResult>>& guesses);             // sqp/utils.h:27
}
struct Instance { size_t dimx, dimy; std::unordered_set<Location> obstacles; std::vector<State> goal_states; };  // Instance.h:27-31
struct VmValue { template <class T> T as() const { return T(); } };
struct Vm { VmValue operator[](const char*) const { return VmValue(); } };
int refine_stage(std::string fname_config, std::vector<Path>& solution, Instance& instance,
                 SolutionStatistics& solution_stat, Vm& vm, bool dump_initial_guess, bool dump_corridor,
                 std::string output_prefix, std::string output_file) {
'''


def _block():
    lines = open(os.path.join(REF, "csdo.cc")).read().splitlines()
    start = next(i for i, l in enumerate(lines) if "2.1. deal with inter-vehicle constraints" in l)
    end = max(i for i, l in enumerate(lines) if l.strip() == "return 1;")
    body = "\n".join(lines[start:end + 1])
    assert "SolverDSQP solver(optimize_res, x0_bar,  inter_planes" in body and "InterpolateInitalGuess" in body
    return body


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "sqp")), reason="needs /root/reference")
@pytest.mark.parametrize("gpu_planes", [False, True])
def test_dropin_compiles_in_reference_tree(tmp_path, gpu_planes):
    body = _block()
    if gpu_planes:
        body, n1 = re.subn(r"(?<![:\w])findNeighborPairsByTrustRegion\(", "csdo_b200::findNeighborPairsByTrustRegion(", body)
        body, n2 = re.subn(r"(?<![:\w])calcEqualInterPlanes\(", "csdo_b200::calcEqualInterPlanes(", body)
        assert n1 == 1 and n2 == 1
    src = tmp_path / "ref_tree_compile.cpp"
    src.write_text(PRELUDE + body + "\n}\n")
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-w", f"-I{REF}", f"-I{ROOT}/include",
           f"-I{ROOT}/oracle/ref_build/stub", "-include", "ostream", "-include", "cmath", str(src)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
