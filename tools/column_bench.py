"""Device-resident F+J+H timing of the Pade column kernel over its instances (2..4 levels, single systems; development tool).
QCK_COLUMN_STAGED=0 / QCK_COLUMN_WPC=n select the direct-store variant / the warps per CTA of the block-staged one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from scan_classes import run, rnd  # noqa: E402  (scan_classes runs its own scan on import when executed as a script only)
from qcknot import workloads as wl  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
systems, traj, integ = wl.config("hadamard", T=T)
run("hadamard N=2 nd=2", systems, traj, integ)
for N, nd, t in ((2, 1, T), (3, 2, T // 2), (4, 2, T // 4), (4, 4, T // 4)):
    run(f"random N={N} nd={nd}", *rnd(N, nd, t))
systems, traj, integ = wl.config("ket", T=T)
run("two kets N=2 nd=2 (shared controls: direct variant)", systems, traj, integ)
run("random ket N=3 nd=2", *rnd(3, 2, T, ket=True))
