#!/usr/bin/env python
"""bench.py -- knot-point constraint+Jacobian+Hessian evals/s on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over the whole (per-GPU) knot range: residual + every Jacobian value + every
Hessian-of-Lagrangian value of every dynamics integrator, written into the solver's fixed-structure value arrays.

  value   device-resident throughput: Z and mu already in HBM, values left in HBM (one fused kernel launch/step)
  e2e     the same metric through the reference-facing host-buffer call (qck_eval_all through QuantumDynamics.eval_all):
          H2D of Z and mu from page-locked host memory + kernel + D2H of F, J, H values, every step
  roofline  algorithmic bytes 8*(2*zdim + 2*dyn + nnzJ + nnzH) per knot block / measured kernel time vs measured HBM peak
  cpu_baseline  oracle/knot_oracle.c (C restatement of the reference algorithm, pthreads over knots) on the host cores

Workload (default): two-transmon (3 levels each, N=9, 4 drives) CZ UnitarySmoothPulseProblem shape, Pade-4 integrator,
free timestep, T = 10,000 knot points per GPU (BASELINE.json north_star target config; configs[3] sweep point),
synthetic random-pulse trajectory, seed 1234.  With N GPUs the knot range is sharded (one-knot halo), T = N*9,999+1:
weak scaling, no data-path collective.  Outputs per step are ~680 MB per GPU, > the 126 MB L2, so no L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FP64_PEAK_TFLOPS = 35.4  # DFMA loop measured on this pool's B200 (profiles/r01_box_peaks_fp64_hbm_pcie.txt)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cz", choices=["cz", "hadamard", "sampling", "ket"])
    ap.add_argument("--integrator", default="pade", choices=["pade", "exponential"])
    ap.add_argument("--T", type=int, default=10000, help="knot points per GPU")
    ap.add_argument("--systems", type=int, default=None, help="sampled systems (sampling workload)")
    ap.add_argument("--shard", default="knot", choices=["knot", "ensemble"],
                    help="knot: every GPU evaluates T knots (weak scaling); ensemble (sampling workload): the sampled systems are "
                         "split over the GPUs, shared-control Hessian entries are summed with one NCCL all-reduce (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline budget (bounded sample)")
    return ap.parse_args()


def workload_name(args, n):
    d = {"cz": "two-transmon(3-level) CZ UnitarySmoothPulseProblem, N=9, 4 drives",
         "hadamard": "single-qubit Hadamard UnitarySmoothPulseProblem, N=2, 2 drives",
         "sampling": f"UnitarySamplingProblem 4-level transmon, {args.systems or 256} systems",
         "ket": "QuantumStateSmoothPulseProblem, N=2, 2 drives"}[args.workload]
    if getattr(args, "shard", "knot") == "ensemble":
        return f"{d}, {args.integrator} integrator, free dt, T={args.T} knots, systems split over {n} GPU(s) (ensemble-sharded)"
    return f"{d}, {args.integrator} integrator, free dt, T={args.T} knots/GPU x {n} GPU(s), knot-sharded"


def algorithmic_bytes(zdim, dyn, nnzJ, nnzH):
    """compulsory reads of z_t, z_t+1, mu_t and writes of F, J, H values per knot block."""
    return 8 * (2 * zdim + 2 * dyn + nnzJ + nnzH)


def flops_per_eval(N, nd, nc):
    """Dense complex products of the implemented Pade-4 algorithm: stage 1: A^2 (N^3), A S, A^H M (N^2 nc each),
    D M^H, S M^H (N^2 nc each); stage 2: A^2 D, (A^2)^H M, C_j D, C_j^H M ((2 + 2 nd) N^2 nc).  8 flops per complex MAC."""
    return 8 * (N ** 3 + (6 + 2 * nd) * N * N * nc)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def build_problem(args, n_gpus):
    from qcknot import workloads as wl
    T_total = n_gpus * (args.T - 1) + 1
    return wl.config(args.workload, T=T_total, integrator=args.integrator, n_systems=args.systems)


def run_reference(args):
    """--impl reference: the reference algorithm's CPU restatement (oracle/knot_oracle.c) on all host cores.
    The Julia reference itself cannot run here or on the box (no julia, its arithmetic lives in un-vendored packages)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.bridge import oracle_dynamics
    from oracle.c_port import CPort
    from qcknot import workloads as wl
    a2 = argparse.Namespace(**vars(args))
    systems, traj, integrators = build_problem(a2, 1)  # one GPU's worth of knots is the bounded sample per step
    O = oracle_dynamics(integrators, traj)
    cp = CPort(O)
    Z = traj.datavec
    nb = traj.T - 1
    mu = wl.random_multipliers(nb * O.dyn)
    cores = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        cp.eval(Z, mu, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cp.eval(Z, mu, nthreads=cores)
    dt = time.perf_counter() - t0
    value = nb * args.steps / dt
    sample = f"{nb} knot blocks/step x {args.steps} steps of the same workload, all {cores} host cores"
    print(json.dumps({
        "impl": "reference", "metric": "knot-pt constraint+Jac+Hess evals/s", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus)},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(args, integrators_builder):
    from oracle.bridge import oracle_dynamics
    from oracle.c_port import CPort
    from qcknot import workloads as wl
    a2 = argparse.Namespace(**vars(args))
    a2.T = min(args.T, 4001)
    systems, traj, integrators = build_problem(a2, 1)
    O = oracle_dynamics(integrators, traj)
    cp = CPort(O)
    Z, nb = traj.datavec, traj.T - 1
    mu = wl.random_multipliers(nb * O.dyn)
    cores = os.cpu_count() or 1
    cp.eval(Z, mu, nthreads=cores)
    reps, t0 = 0, time.perf_counter()
    while True:
        cp.eval(Z, mu, nthreads=cores)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > args.cpu_seconds or reps >= 200:
            break
    return {"value": nb * reps / dt, "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": f"{nb} knot blocks of the same workload x {reps} passes ({dt:.1f} s), pthreads over knots"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import qcknot
    from qcknot import workloads as wl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    n_gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")

    ensemble = args.shard == "ensemble"
    if ensemble and args.workload != "sampling":
        raise SystemExit("--shard ensemble needs --workload sampling")
    if ensemble:
        # strong scaling: the same T knots on every GPU, each GPU owns a slice of the sampled systems
        from qcknot.sharding import integrator_shard
        systems, traj, integrators = build_problem(args, 1)
        q0, q1 = integrator_shard(len(systems), len(integrators), rank, n_gpus)
        D = qcknot.QuantumDynamics(integrators, traj, device=local_rank, integrator_range=(q0, q1))
        nb = D.n_blocks
        Zh = np.ascontiguousarray(traj.datavec[: (nb + 1) * D.zdim])
        muh = wl.random_multipliers(nb * D.dyn, seed=1234)
    else:
        systems, traj, integrators = build_problem(args, n_gpus)
        nbp = args.T - 1
        t0k, t1k = rank * nbp, (rank + 1) * nbp
        D = qcknot.QuantumDynamics(integrators, traj, device=local_rank, knot_range=(t0k, t1k))
        nb = D.n_blocks
        Zh = np.ascontiguousarray(traj.datavec[t0k * D.zdim:(t1k + 1) * D.zdim])
        muh = wl.random_multipliers(nb * D.dyn, seed=1234 + rank)

    # ---- device-resident arm ------------------------------------------------------------------------------------------
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    Z = torch.from_numpy(Zh).to(dev)
    mu = torch.from_numpy(muh).to(dev)
    F = torch.empty(nb * D.dyn, dtype=torch.float64, device=dev)
    J = torch.empty(nb * D.nnzJ, dtype=torch.float64, device=dev)
    H = torch.empty(nb * max(D.nnzH, 1), dtype=torch.float64, device=dev)
    st = stream.cuda_stream

    shared_idx = None
    if ensemble and world > 1:
        pos = torch.from_numpy(D.shared_hessian_positions()).to(dev)
        shared_idx = (torch.arange(nb, device=dev)[:, None] * D.nnzH + pos[None, :]).reshape(-1)

    def step():
        D.eval_device(7, Z.data_ptr(), mu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), st)
        if shared_idx is not None:  # the one data-path collective: partial sums of the entries on the shared controls
            buf = H[shared_idx]
            dist.all_reduce(buf)
            H[shared_idx] = buf

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = D.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = D.launch_count - l0

    # ---- end-to-end arm: host buffers through the reference-facing call --------------------------------------------------
    Fh, Jh, Hh = np.empty(nb * D.dyn), np.empty(nb * D.nnzJ), np.empty(nb * max(D.nnzH, 1))
    for a in (Zh, muh, Fh, Jh, Hh):
        qcknot.host_register(a)
    e2e_steps = max(3, min(args.steps, 10))
    D.eval_all(Zh, muh, Fh, Jh, Hh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        D.eval_all(Zh, muh, Fh, Jh, Hh)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    checksum = float(Fh.sum() + Jh[:: 997].sum() + Hh[:: 997].sum())

    tms = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), float(tms[1])
    total_evals_per_step = nb if ensemble else nb * n_gpus
    value = total_evals_per_step * args.steps / (ms * 1e-3)
    e2e_value = total_evals_per_step * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        bpe = algorithmic_bytes(D.zdim, D.dyn, D.nnzJ, D.nnzH)
        kernel_s = ms * 1e-3 / args.steps
        achieved = bpe * nb / kernel_s * 1e-9 / (n_gpus if ensemble else 1)  # per GPU
        peak, peak_src = 6650.0, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(f"{args.workload}_{args.integrator}_T{args.T}")
        except Exception:
            pass
        q = next(I for I in integrators if hasattr(I, "system"))
        nq = sum(1 for I in integrators if hasattr(I, "system"))
        fl = flops_per_eval(q.system.levels, q.system.n_drives, q.system.levels if q.unitary else 1) * nq
        out = {
            "metric": "knot-pt constraint+Jac+Hess evals/s", "value": value, "unit": "evals/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if ensemble else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, n_gpus), "seed": 1234, "evals_per_step": total_evals_per_step,
                       "l2": "outputs per step exceed L2 (inputs+outputs larger than L2, no flush needed)"
                       if bpe * nb > 2 * 126e6 else "working set fits L2; flush not applied"},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(Zh.nbytes + muh.nbytes),
                    "d2h_bytes_per_step": int(Fh.nbytes + Jh.nbytes + Hh.nbytes), "steps": e2e_steps,
                    "timing": "host wall clock around synchronous qck_eval_all calls, max over ranks"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)",
                         "bytes_per_eval": bpe, "kernel_us": kernel_s * 1e6,
                         "fp64": {"flops_per_eval": fl, "achieved_tflops": fl * nb / kernel_s * 1e-12,
                                  "peak_tflops": FP64_PEAK_TFLOPS,
                                  "frac": fl * nb / kernel_s * 1e-12 / FP64_PEAK_TFLOPS}},
            "checksum": checksum,
        }
        if n_gpus == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, None)
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out))
    for a in (Zh, muh, Fh, Jh, Hh):
        qcknot.host_unregister(a)
    D.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
