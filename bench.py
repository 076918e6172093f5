#!/usr/bin/env python
"""bench.py -- knot-point constraint+Jacobian+Hessian evals/s on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over the whole (per-GPU) knot range: residual + every Jacobian value + every
Hessian-of-Lagrangian value of every dynamics integrator, written into the solver's fixed-structure value arrays.

  value     device-resident throughput: Z and mu already in HBM, values left in HBM (one fused kernel launch per step).
            The timed region is stretched to >= 1 s by repeating the K steps R times (config.timed_passes = K * R); all
            numbers are per step.
  e2e       the same metric through the reference-facing host-buffer call (QuantumDynamics.eval_all -> qck_eval_all) with
            ordinary (pageable) numpy arrays, every step on a DIFFERENT trajectory (two alternate) so that nothing is cached:
            staging + H2D of Z and mu, kernels, compact D2H (kron blocks once), host-side expansion into the caller's arrays.
            With N > 1 GPUs ONE caller (rank 0) drives all N GPUs through one multi-GPU handle (n_gpus = N, knot-sharded);
            the other ranks idle behind a host-side barrier.
  roofline  algorithmic bytes 8*(2*zdim + 2*dyn + nnzJ + nnzH) per knot block / measured kernel time vs measured HBM peak;
            `pcie`: bytes the e2e step moved over PCIe / e2e time vs the measured link; `host`: bytes written into the
            caller's arrays / e2e time next to what the host threads reach on their own (expansion only, no GPU involved).
  cpu_baseline  oracle/knot_oracle.c (C restatement of the reference algorithm, pthreads over knots) on the host cores:
            `value` = the tuned variant (constant anticommutators hoisted, G_j G + G G_j reused), `literal` = the naive one.
  extra     sub-records (rank 0): exponential integrators, T = 100,000, the sampling problem (ensemble-sharded when N > 1),
            the NCCL all-gather of the device-resident segments (N > 1).

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
PCIE_PEAK_GBS = 56.0     # pinned D2H measured on this pool's B200 boxes (same file)
METRIC = "knot-pt constraint+Jac+Hess evals/s"


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra sub-records")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU-baseline budget (bounded sample)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="minimum length of the timed regions")
    return ap.parse_args()


def workload_name(args, n):
    d = {"cz": "two-transmon(3-level) CZ UnitarySmoothPulseProblem, N=9, 4 drives",
         "hadamard": "single-qubit Hadamard UnitarySmoothPulseProblem, N=2, 2 drives",
         "sampling": f"UnitarySamplingProblem 4-level transmon, {args.systems or 256} systems",
         "ket": "QuantumStateSmoothPulseProblem, N=2, 2 drives"}[args.workload]
    return f"{d}, {args.integrator} integrator, free dt, T={args.T} knots/GPU x {n} GPU(s), knot-sharded"


def config_dict(args, n):
    """Identical in both arms (the driver compares them)."""
    return {"workload": workload_name(args, n), "seed": 1234, "evals_per_step": n * (args.T - 1),
            "l2": "outputs per step exceed L2 (inputs+outputs larger than L2, no flush needed)"}


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


def build_problem(args, n_gpus, **kw):
    from qcknot import workloads as wl
    T_total = n_gpus * (args.T - 1) + 1
    return wl.config(kw.get("workload", args.workload), T=kw.get("T", T_total), integrator=kw.get("integrator", args.integrator),
                     n_systems=kw.get("systems", args.systems))


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU legs (the only places that execute oracle/) --------------------------------------------------------------------------
def _cpu_port(args, T):
    from oracle.bridge import oracle_dynamics
    from oracle.c_port import CPort
    from qcknot import workloads as wl
    a2 = argparse.Namespace(**vars(args))
    a2.T = T
    systems, traj, integrators = build_problem(a2, 1)
    O = oracle_dynamics(integrators, traj)
    cp = CPort(O)
    nb = traj.T - 1
    return cp, traj.datavec, wl.random_multipliers(nb * O.dyn), nb


def run_reference(args):
    """--impl reference: the reference algorithm's CPU restatement (oracle/knot_oracle.c, tuned variant) on all host cores.
    The Julia reference itself cannot run here or on the box (no julia, its arithmetic lives in un-vendored packages)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cp, Z, mu, nb = _cpu_port(args, args.T)  # one GPU's worth of knots is the bounded sample per step
    cores = os.cpu_count() or 1
    out = cp.eval(Z, mu, nthreads=cores, tuned=True)
    for _ in range(min(args.warmup, 2)):
        cp.eval(Z, mu, nthreads=cores, tuned=True, out=out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cp.eval(Z, mu, nthreads=cores, tuned=True, out=out)
    dt = time.perf_counter() - t0
    value = nb * args.steps / dt
    sample = (f"{nb} knot blocks/step x {args.steps} steps of the same workload, all {cores} host cores, tuned variant "
              "(constant anticommutators hoisted, G_j G + G G_j reused)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(args):
    cp, Z, mu, nb = _cpu_port(args, min(args.T, 4001))
    cores = os.cpu_count() or 1
    res = {}
    for tuned, budget in ((True, 0.7 * args.cpu_seconds), (False, 0.3 * args.cpu_seconds)):
        out = cp.eval(Z, mu, nthreads=cores, tuned=tuned)
        reps, t0 = 0, time.perf_counter()
        while True:
            cp.eval(Z, mu, nthreads=cores, tuned=tuned, out=out)
            reps += 1
            dt = time.perf_counter() - t0
            if dt > budget or reps >= 200:
                break
        res[tuned] = (nb * reps / dt, reps, dt)
    return {"value": res[True][0], "unit": "evals/s", "cores": cores, "kind": "port", "literal": res[False][0],
            "sample": f"{nb} knot blocks of the same workload x {res[True][1]} passes ({res[True][2]:.1f} s) tuned "
                      f"(anticommutators hoisted, G_j G + G G_j reused) / x {res[False][1]} passes literal, pthreads over knots"}


# ---- device-resident timing of one handle ------------------------------------------------------------------------------------
def time_device(D, Z, mu, F, J, H, stream, steps, warmup, min_seconds, barrier=None, mask=7):
    """Returns (ms per pass, passes timed, launches in the timed region)."""
    import torch
    st = stream.cuda_stream

    def step():
        D.eval_device(mask, Z.data_ptr(), mu.data_ptr(), F.data_ptr(), J.data_ptr(), H.data_ptr(), st)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(max(warmup, 3)):
        step()
    e0.record(stream)
    for _ in range(3):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1) / 3, 1e-3)
    reps = max(1, int(np.ceil(min_seconds * 1e3 / (steps * est))))
    if barrier:
        reps = barrier(reps)  # the same repeat count on every rank
    l0 = D.launch_count
    if barrier:
        barrier(0)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps * reps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if barrier:
        barrier(0)
    return e0.elapsed_time(e1) / (steps * reps), steps * reps, D.launch_count - l0


def time_e2e(D, Zs, mus, F, J, H, min_seconds, min_steps=5, max_steps=400):
    """Host-buffer calls on alternating trajectories; returns (s per step, steps, transfer stats of the last step)."""
    for i in range(3):
        D.eval_all(Zs[i & 1], mus[i & 1], F, J, H)
    t0 = time.perf_counter()
    D.eval_all(Zs[1], mus[1], F, J, H)
    est = time.perf_counter() - t0
    steps = int(min(max_steps, max(min_steps, np.ceil(min_seconds / est))))
    t0 = time.perf_counter()
    for i in range(steps):
        D.eval_all(Zs[i & 1], mus[i & 1], F, J, H)
    dt = (time.perf_counter() - t0) / steps
    return dt, steps, D.transfer_stats()


def pcie_peak_gbs(dev, mb=256):
    """Pinned D2H / H2D copy rate of this box's link, measured live (GB/s each way): the denominator of the pcie sub-roofline.
    CUDA events on a stream of its own, two untimed copies first, best of six."""
    import torch
    n = mb * (1 << 20) // 8
    hbuf = torch.empty(n, dtype=torch.float64).pin_memory()
    dbuf = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.cuda.Stream(device=dev)
    out = {}
    with torch.cuda.device(dev), torch.cuda.stream(st):
        for name, (dst, src) in (("d2h", (hbuf, dbuf)), ("h2d", (dbuf, hbuf))):
            best = 0.0
            for rep in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(dev)
                e0.record(st)
                dst.copy_(src, non_blocking=True)
                e1.record(st)
                e1.synchronize()
                if rep >= 2:
                    best = max(best, n * 8 / (e0.elapsed_time(e1) * 1e-3) * 1e-9)
            out[name] = best
    return out


def expand_only_gbs(D, nb):
    """The host half alone (compact layout -> caller arrays, no GPU): what the host threads and memory system can absorb."""
    bufs, out_bytes = [], 0
    for arr, nnz in ((0, D.dyn), (1, D.nnzJ), (2, D.nnzH)):
        C_ = int(D.compact_map(arr)[:, 2].sum())
        bufs.append((arr, np.zeros(nb * C_), np.empty(nb * nnz)))
        out_bytes += nb * nnz * 8
    best = 0.0
    for rep in range(4):
        t0 = time.perf_counter()
        for arr, comp, out in bufs:
            D.expand_host(arr, comp, out, nb)
        if rep:
            best = max(best, out_bytes / (time.perf_counter() - t0) * 1e-9)
    return best


def extra_records(args, dev, stream, n_gpus, rank0_devices):
    """Sub-records beside the headline (rank 0 only): every one is a device-timed pass through the same C-ABI."""
    import torch
    import qcknot
    from qcknot import workloads as wl
    peak, _ = hbm_peak()
    out = {}

    def dev_arm(workload, T, integrator, systems=None, min_seconds=0.3):
        a2 = argparse.Namespace(**vars(args))
        a2.workload, a2.T, a2.integrator, a2.systems = workload, T, integrator, systems
        systems_, traj, integrators = build_problem(a2, 1)
        D = qcknot.QuantumDynamics(integrators, traj, device=dev.index)
        nb = D.n_blocks
        Z = torch.from_numpy(traj.datavec[: traj.T * D.zdim]).to(dev)
        mu = torch.from_numpy(wl.random_multipliers(nb * D.dyn)).to(dev)
        F = torch.empty(nb * D.dyn, dtype=torch.float64, device=dev)
        J = torch.empty(nb * D.nnzJ, dtype=torch.float64, device=dev)
        H = torch.empty(nb * max(D.nnzH, 1), dtype=torch.float64, device=dev)
        ms, passes, launches = time_device(D, Z, mu, F, J, H, stream, 5, 3, min_seconds)
        bpe = algorithmic_bytes(D.zdim, D.dyn, D.nnzJ, D.nnzH)
        rec = {"evals_per_s": nb / (ms * 1e-3), "us_per_pass": ms * 1e3, "knot_blocks": nb, "passes": passes,
               "launches_per_pass": launches / passes, "bytes_per_eval": bpe, "hbm_gbs": bpe * nb / (ms * 1e-3) * 1e-9,
               "hbm_frac": bpe * nb / (ms * 1e-3) * 1e-9 / peak}
        D.close()
        del Z, mu, F, J, H
        return rec, (systems_, traj, integrators)

    # the residual the north_star names: U_{t+1} - exp(-i H dt) U_t (F + J + H, the Hessian is ours: the reference has none)
    rec, _ = dev_arm("cz", args.T, "exponential")
    # spectral algorithm (csrc/qck_expeig.cu): Jacobi eigen-decomposition + divided differences, ~0.30 MFLOP per
    # 9-level knot (31 product-sized steps of 8 N^3 flop, the second-order contraction, ~27 Jacobi rounds); two launches per pass
    flops = 0.30e6
    rec["fp64"] = {"flops_per_eval": flops, "achieved_tflops": flops * rec["evals_per_s"] * 1e-12, "peak_tflops": FP64_PEAK_TFLOPS,
                   "frac": flops * rec["evals_per_s"] * 1e-12 / FP64_PEAK_TFLOPS}
    rec["note"] = ("F+J+H through the eigen + spectral kernels; roofline = min(HBM, FP64): neither binds (dependent latency at 8 warps "
                   "per SM); the scaling-and-squaring kernel this replaces needs 1.48 MFLOP per knot and ran at 4.7 M evals/s")
    out["exponential"] = rec
    # the same residual on the small configurations (spectral column kernels, csrc/qck_colexp.cu)
    rec, _ = dev_arm("hadamard", 100000, "exponential")
    out["hadamard_exponential_T100000"] = rec
    rec, _ = dev_arm("sampling", 200, "exponential", systems=256)
    out["sampling_exponential_256_T200"] = rec
    rec, _ = dev_arm("cz", 100000, "pade")
    out["cz_T100000"] = rec
    rec, _ = dev_arm("sampling", 200, "pade", systems=256)
    out["sampling_256_T200"] = rec
    rec, _ = dev_arm("hadamard", 100000, "pade")
    out["hadamard_T100000"] = rec
    return out


def extra_ensemble(args, n_gpus):
    """BASELINE configs[4]: UnitarySamplingProblem, 256 sampled systems on a 4-level transmon, the ensemble split over the GPUs of
    ONE handle (strong scaling: same problem, more GPUs).  Device-resident: kernels + the sum of the shared-control Hessian
    entries over peer memory; e2e: host buffers through qck_eval_all."""
    import qcknot
    from qcknot import workloads as wl
    out = {}
    for T in (200, 2000):
        a2 = argparse.Namespace(**vars(args))
        a2.workload, a2.T, a2.integrator, a2.systems = "sampling", T, "pade", 256
        systems, traj, integrators = build_problem(a2, 1)
        D = qcknot.QuantumDynamics(integrators, traj, device=0, n_gpus=n_gpus, shard_mode="ensemble") if n_gpus > 1 else \
            qcknot.QuantumDynamics(integrators, traj, device=0)
        nb = D.n_blocks
        Z = np.ascontiguousarray(traj.datavec[: traj.T * D.zdim])
        mu = wl.random_multipliers(nb * D.dyn)
        D.upload(Z, mu)
        for _ in range(3):
            D.eval_resident(7)
        D.synchronize()
        reps = 20 if T == 200 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            D.eval_resident(7)
        D.synchronize()
        dev_s = (time.perf_counter() - t0) / reps
        Fh, Jh, Hh = np.empty(nb * D.dyn), np.empty(nb * D.nnzJ), np.empty(nb * D.nnzH)
        Zs, mus = [Z, Z + 1e-7], [mu, wl.random_multipliers(nb * D.dyn, seed=7)]
        e2e_s, steps, st = time_e2e(D, Zs, mus, Fh, Jh, Hh, 0.3, min_steps=3, max_steps=40)
        out[f"T{T}"] = {"knot_blocks": nb, "systems": 256, "n_gpus": n_gpus,
                        "device_us_per_pass": dev_s * 1e6, "device_knot_evals_per_s": nb / dev_s,
                        "device_timing": "host wall clock around eval_resident x reps + synchronize (one caller, all GPUs)",
                        "e2e_ms_per_step": e2e_s * 1e3, "e2e_knot_evals_per_s": nb / e2e_s, "e2e_steps": steps,
                        "d2h_bytes_per_step": int(st["d2h_bytes"])}
        D.close()
    return out


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
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        host_group = dist.new_group(backend="gloo")  # host-side barrier: does not occupy the GPUs while rank 0 drives them
    n_gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")

    # ---- device-resident arm: one process per GPU, rank g evaluates its knot shard (weak scaling, no data-path collective) ----
    systems, traj, integrators = build_problem(args, n_gpus)
    nbp = args.T - 1
    t0k, t1k = rank * nbp, (rank + 1) * nbp
    D = qcknot.QuantumDynamics(integrators, traj, device=local_rank, knot_range=(t0k, t1k))
    nb = D.n_blocks
    Zh = np.ascontiguousarray(traj.datavec[t0k * D.zdim:(t1k + 1) * D.zdim])
    muh = wl.random_multipliers(nb * D.dyn, seed=1234 + rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    Z = torch.from_numpy(Zh).to(dev)
    mu = torch.from_numpy(muh).to(dev)
    F = torch.empty(nb * D.dyn, dtype=torch.float64, device=dev)
    J = torch.empty(nb * D.nnzJ, dtype=torch.float64, device=dev)
    H = torch.empty(nb * max(D.nnzH, 1), dtype=torch.float64, device=dev)

    def barrier(reps):
        torch.cuda.synchronize()
        if world > 1:
            t = torch.tensor([reps], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            reps = int(t.item())
        torch.cuda.synchronize()
        return reps

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, passes, launches = time_device(D, Z, mu, F, J, H, stream, args.steps, args.warmup, args.min_seconds, barrier)
    clocks = sampler.stop()
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms[0])
    checksum = float(F.sum().item() + J[::997].sum().item() + H[::997].sum().item())
    del Z, mu, F, J, H
    torch.cuda.empty_cache()

    # ---- end-to-end arm: ONE caller, host buffers, through the reference-facing call -----------------------------------------------
    e2e = None
    host = None
    if rank == 0:
        if n_gpus == 1:
            De, nbe = D, nb
            Ze = [Zh, Zh + 1e-7]
            mue = [muh, wl.random_multipliers(nb * D.dyn, seed=99)]
        else:
            De = qcknot.QuantumDynamics(integrators, traj, device=0, n_gpus=n_gpus, shard_mode="knot")
            nbe = De.n_blocks
            Z0 = np.ascontiguousarray(traj.datavec[: traj.T * D.zdim])
            Ze = [Z0, Z0 + 1e-7]
            mue = [wl.random_multipliers(nbe * D.dyn, seed=1234), wl.random_multipliers(nbe * D.dyn, seed=99)]
        Fh, Jh, Hh = np.empty(nbe * D.dyn), np.empty(nbe * D.nnzJ), np.empty(nbe * max(D.nnzH, 1))
        link = pcie_peak_gbs(dev)
        # pageable caller arrays first (everything staged by the library), then the same arrays page-locked with qck_host_register
        # (what the Julia shim does once with Ipopt's value buffers): F and the Hessian values then leave by DMA
        # straight to their final place; the headline is the registered run
        pg_s, pg_steps, _ = time_e2e(De, Ze, mue, Fh, Jh, Hh, 0.5 * args.min_seconds)
        for arr in (Fh, Jh, Hh):
            qcknot.host_register(arr)
        e2e_s, e2e_steps, st = time_e2e(De, Ze, mue, Fh, Jh, Hh, args.min_seconds)
        for arr in (Fh, Jh, Hh):
            qcknot.host_unregister(arr)
        moved = st["h2d_bytes"] + st["d2h_bytes"]
        written = Fh.nbytes + Jh.nbytes + Hh.nbytes
        e2e = {"value": nbe / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(st["h2d_bytes"]),
               "d2h_bytes_per_step": int(st["d2h_bytes"]), "steps": e2e_steps, "ms_per_step": e2e_s * 1e3,
               "caller": "one host thread calling qck_eval_all" + (f" on one handle with n_gpus={n_gpus} (knot-sharded)" if n_gpus > 1 else ""),
               "inputs": "pageable numpy arrays for Z and mu, two alternating trajectories (no cache hits); output arrays page-locked once "
                         "with qck_host_register",
               "pageable_outputs": {"value": nbe / pg_s, "ms_per_step": pg_s * 1e3, "steps": pg_steps},
               "timing": "host wall clock around synchronous calls",
               "pcie": {"bytes_per_step": int(moved), "achieved_gbs": moved / e2e_s * 1e-9, "links": n_gpus,
                        "d2h_achieved_gbs": st["d2h_bytes"] / e2e_s * 1e-9,
                        "peak_gbs": link["d2h"] * n_gpus, "frac": st["d2h_bytes"] / e2e_s * 1e-9 / (link["d2h"] * n_gpus),
                        "peak_source": f"pinned 256 MB copies measured in this run on GPU 0: D2H {link['d2h']:.1f} GB/s, H2D {link['h2d']:.1f} GB/s "
                                       f"(round-1 box: {PCIE_PEAK_GBS} GB/s); frac = D2H bytes / e2e time / D2H peak (H2D runs the other way, concurrently)",
                        "full_layout_bytes": int(written + Ze[0].nbytes + mue[0].nbytes)}}
        exp_gbs = expand_only_gbs(De, min(nbe, 9999))
        host = {"written_bytes_per_step": int(written), "written_gbs": written / e2e_s * 1e-9, "expand_only_gbs": exp_gbs,
                "frac_of_expand_only": written / e2e_s * 1e-9 / exp_gbs if exp_gbs else None, "threads": os.cpu_count(),
                "note": "expand_only = the library's host threads writing the caller's arrays from the compact layout with no GPU "
                        "involved: the ceiling of the e2e path on this host's memory system"}
        checksum += float(Fh.sum() + Jh[::997].sum() + Hh[::997].sum())
        if De is not D:
            extra_multi = {}
            try:
                import ctypes
                ver, nranks = De.nccl_version()
                De.upload(Ze[0], mue[0])
                for _ in range(2):
                    De.eval_resident(7)
                    De.gather_device(7)
                De.synchronize()
                t0 = time.perf_counter()
                reps = 5
                for _ in range(reps):
                    De.eval_resident(7)
                    De.gather_device(7)
                De.synchronize()
                dtg = (time.perf_counter() - t0) / reps
                t0 = time.perf_counter()
                for _ in range(reps):
                    De.eval_resident(7)
                De.synchronize()
                dte = (time.perf_counter() - t0) / reps
                per_gpu_bytes = written * (n_gpus - 1) / n_gpus
                extra_multi = {"nccl_version": ver, "nccl_nranks": nranks, "eval_resident_ms": dte * 1e3,
                               "eval_plus_allgather_ms": dtg * 1e3,
                               "allgather_gbs_per_gpu_in": per_gpu_bytes / max(dtg - dte, 1e-9) * 1e-9,
                               "note": "one caller: kernels on every GPU, then ncclBroadcast-grouped all-gather of the F/J/H segments "
                                       "so that every GPU holds the assembled arrays"}
            except Exception as ex:  # noqa: BLE001
                extra_multi = {"error": str(ex)}
            e2e["gather"] = extra_multi
            De.close()
        if not args.no_extra:
            try:
                e2e_ens = extra_ensemble(args, n_gpus)
            except Exception as ex:  # noqa: BLE001
                e2e_ens = {"error": str(ex)}
        else:
            e2e_ens = None
    if host_group is not None:
        dist.barrier(group=host_group)

    extra = None
    if rank == 0 and not args.no_extra:
        try:
            extra = extra_records(args, dev, stream, n_gpus, None)
        except Exception as ex:  # noqa: BLE001
            extra = {"error": str(ex)}
        extra["ensemble_sampling_256"] = e2e_ens
    if host_group is not None:
        dist.barrier(group=host_group)

    if rank == 0:
        bpe = algorithmic_bytes(D.zdim, D.dyn, D.nnzJ, D.nnzH)
        kernel_s = ms * 1e-3
        achieved = bpe * nb / kernel_s * 1e-9  # per GPU
        peak, peak_src = hbm_peak()
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(f"{args.workload}_{args.integrator}_T{args.T}")
        except Exception:
            pass
        q = next(I for I in integrators if hasattr(I, "system"))
        nq = sum(1 for I in integrators if hasattr(I, "system"))
        fl = flops_per_eval(q.system.levels, q.system.n_drives, q.system.levels if q.unitary else 1) * nq
        cfg = config_dict(args, n_gpus)
        cfg_run = {"timed_passes": passes, "timed_seconds": ms * 1e-3 * passes}
        out = {
            "metric": METRIC, "value": nb * n_gpus / kernel_s, "unit": "evals/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "run": cfg_run,
            "e2e": e2e, "host": host,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_eval": bpe, "kernel_us": kernel_s * 1e6,
                         "fp64": {"flops_per_eval": fl, "achieved_tflops": fl * nb / kernel_s * 1e-12,
                                  "peak_tflops": FP64_PEAK_TFLOPS,
                                  "frac": fl * nb / kernel_s * 1e-12 / FP64_PEAK_TFLOPS}},
            "checksum": checksum,
            "extra": extra,
        }
        if n_gpus == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args)
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out))
    D.close()
    if world > 1:
        dist.barrier(group=host_group)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
