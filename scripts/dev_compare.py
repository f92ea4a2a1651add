"""Developer check (test infrastructure): GPU refine vs CPU oracle on small synthetic batches."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.scenario import synthetic_instance
from csdotrajectoryplanning_b200.solver import DsqpSolver
from oracle import oracle as O

p = default_params()
args = dict(a.split("=") for a in sys.argv[1:])
for k in ("osqp_max_iter", "max_iter"):
    if k in args: setattr(p, k, int(args[k]))
n_inst = int(args.get("n_inst", 2)); na = int(args.get("na", 6)); no = int(args.get("no", 25))
size = float(args.get("size", 50)); acts = int(args.get("acts", 20))
inst = [synthetic_instance(100 + i, size, na, no, (acts // 2, acts)) for i in range(n_inst)]
b0 = pack_instances(inst)
S = DsqpSolver(p)
# planes
pb, legal = S.planes(b0)
ok = True
for i, ins in enumerate(inst):
    pts, pabc, lg = O.instance_planes(p, ins.guess)
    ins.plane_t, ins.plane_abc = pts, pabc
b = pack_instances(inst)
print("planes: ptr equal", np.array_equal(pb.plane_ptr, b.plane_ptr), "t equal", np.array_equal(pb.plane_t, b.plane_t),
      "abc maxdiff", np.abs(pb.plane_abc - b.plane_abc).max() if b.plane_abc.size else 0, "bit-equal", np.array_equal(pb.plane_abc, b.plane_abc), "K", int(b.plane_ptr[-1]))
# corridors
corr, bs, lg = S.corridors(b, False)
mx = 0; nbad = 0
for a in range(b.n_agents):
    i = int(np.searchsorted(b.inst_agent_ptr, a, side="right") - 1)
    g = b.agent_guess(a); nt = g.shape[1]; o = int(b.agent_off[a])
    ob = b.obs[3*b.obs_ptr[i]:3*b.obs_ptr[i+1]].reshape(-1, 3)
    c0, s0, l0 = O.agent_corridors(p, g[0], g[1], g[2], b.inst_dims[2*i], b.inst_dims[2*i+1], ob, False)
    c1 = corr[8*o:8*(o+nt)].reshape(8, nt)
    nbad += int((c0 != c1).sum()); mx = max(mx, np.abs(c0 - c1).max())
print("corridors: mismatching entries", nbad, "maxdiff", mx, "legal", lg)
t = time.time(); r_o, fl = O.refine(p, b, linsys=0, nthreads=8); t_o = time.time() - t
t = time.time(); r_g = S.refine(b); t_g = time.time() - t
t = time.time(); r_g = S.refine(b); t_g2 = time.time() - t
print("launch", S.last_launch())
print("oracle time %.3f s, gpu %.3f s (2nd %.3f s), agents %d, qps %d" % (t_o, t_g, t_g2, b.n_agents, r_o.sqp_iters.sum()))
print("status  o", r_o.status, "\n        g", r_g.status)
print("sqp     o", r_o.sqp_iters, "\n        g", r_g.sqp_iters)
print("admm    o", r_o.admm_iters, "\n        g", r_g.admm_iters)
print("nfac    o", r_o.n_factor, "\n        g", r_g.n_factor)
print("inst_status", r_o.inst_status, r_g.inst_status, "legal", r_o.inst_static_legal, r_g.inst_static_legal)
d = np.abs(r_o.traj - r_g.traj)
print("traj maxdiff", d.max(), "corr maxdiff", np.abs(r_o.corridors - r_g.corridors).max(), "obj maxdiff", np.abs(r_o.objective - r_g.objective).max())
for a in range(b.n_agents):
    o = int(b.agent_off[a]); nt = int(b.agent_off[a+1]) - o
    da = d[6*o:6*(o+nt)].reshape(6, nt).max(axis=1)
    print(" agent", a, "K", int(b.plane_ptr[a+1]-b.plane_ptr[a]), "maxdiff per var", np.array2string(da, precision=2))
