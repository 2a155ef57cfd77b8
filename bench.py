#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 engine for the ElPhDynamics hot path.

Metric (BASELINE.json): M^T M matvecs/s on Holstein square 32x32, Ltau=200 (config B), plus
Langevin steps/s and CG iterations/s as extra keys, with the kernel's fraction of the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the fused M^T M kernel over one batch of
synthetic input: R independent replicas of the 32x32x200 lattice (own phonon field / expnV table and own
vector each -- the reference's only scale-out is independent runs, src/ElPhDynamics.jl:90-95), sized so
that the inputs (v + expnV + y = R * 4.9 MB) exceed the 126 MB L2.  `value` times the kernel with inputs
resident in HBM; `e2e` times the reference-facing C-ABI call (elph_mulMTM_batch) with pinned HOST buffers,
H2D + layout change + kernel + D2H inside the timed region.  N > 1: replicas are sharded across ranks
(no data-path collective; weak scaling), max-over-ranks timing.

--impl reference: the reference's CPU implementation cannot run here (pure Julia, no julia in the image),
so the arm times the C restatement of its loops (oracle/c/elph_ref.c) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "MTM matvecs/s (Holstein 32x32xL200)"
UNIT = "matvecs/s"
LSIDE, BETA, DTAU = 32, 20.0, 0.1
BYTES_PER_POINT = 24.0  # read v + read expnV + write y


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def bind_to_gpu_numa_node(local: int):
    """Pin this process to the host cores of the NUMA node its GPU hangs off, BEFORE any page-locked staging buffer is
    allocated (first-touch places the pages on that node): with 8 ranks the host<->device copies of the e2e leg otherwise cross
    the socket interconnect.  Returns a description for the bench line, or None when the topology cannot be read."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text().strip())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"pci": bus, "numa_node": node, "cores": len(cpus)}
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_model():
    """Configuration B through the package's own host API (no oracle on the product arm)."""
    from elphdynamics_b200 import workloads
    return workloads.holstein("square", LSIDE, BETA, DTAU, mu=-1.0, seed=1234, eps=0.3)


def oracle_model():
    """The same configuration and field as an oracle object: ONLY for the cpu_baseline / --impl reference legs."""
    from helpers import oracle_holstein
    return oracle_holstein("square", LSIDE, BETA, DTAU, mu=-1.0, seed=1234, eps=0.3)[0]


def cpu_baseline(target_seconds=12.0, nthreads=0):
    """C restatement of the reference loops on the host cores; bounded sample of the same workload."""
    from oracle.cref import CRef
    om = oracle_model()
    c = CRef(om, native=True)
    ncpu = os.cpu_count() or 1
    nthreads = nthreads or ncpu
    nrep = nthreads
    secs, used = c.mulMTM_throughput(nrep=nrep, reps=2, nthreads=nthreads)
    per = secs / 2
    reps = max(2, int(target_seconds / max(per, 1e-6)))
    secs, used = c.mulMTM_throughput(nrep=nrep, reps=reps, nthreads=nthreads)
    return {"value": nrep * reps / secs, "unit": UNIT, "cores": used, "kind": "port",
            "sample": f"{nrep} independent 32x32xL200 replicas x {reps} M^T M products each, one thread per replica "
                      f"(C restatement of the Julia loops, gcc -O3 -march=native -ffast-math); the Julia reference itself "
                      f"cannot run here", "seconds": secs}, (nrep, reps, secs, used)


def cpu_langevin_baseline(nsteps=2):
    """The second headline quantity on the host: Langevin Runge-Kutta steps/s at config B with the shipped solver settings
    (KPM-preconditioned CG).  oracle/cfast.py: the reference's loops restated in C (products, force, tau-averaged operator,
    Chebyshev recurrences), the FFTs through pocketfft and the 20 x 20 eigenvalue problem through LAPACK, as the reference
    hands them to FFTW and LAPACK.  One chain on one thread (the reference's own setting, src/ElPhDynamics.jl:74-75) and one
    chain per core (its own scale-out).  Runs in a child process so that the BLAS / OpenMP pools are pinned to one thread."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "oracle.cfast", "--steps", str(nsteps)], cwd=str(ROOT), env=env, capture_output=True,
                       text=True, timeout=600)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-300:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    om = oracle_model()
    from oracle.cref import CRef
    c = CRef(om, native=True)
    nthreads = os.cpu_count() or 1
    nrep = nthreads
    secs, used = c.mulMTM_throughput(nrep=nrep, reps=2, nthreads=nthreads)
    reps = max(1, int(1.0 / max(secs / 2, 1e-6)))     # ~1 s of CPU work per step
    for _ in range(args.warmup):
        c.mulMTM_throughput(nrep=nrep, reps=reps, nthreads=nthreads)
    t = 0.0
    for _ in range(args.steps):
        s, used = c.mulMTM_throughput(nrep=nrep, reps=reps, nthreads=nthreads)
        t += s
    value = nrep * reps * args.steps / t
    sample = (f"each step = {nrep} independent 32x32xL200 replicas x {reps} M^T M products, one thread per replica; "
              f"C restatement of the reference's Julia loops (Julia itself is not installed in this image)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "holstein_square_32x32_L200", "replicas_per_step": nrep, "products_per_replica": reps},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    try:        # the second headline quantity of the same reference arm: Langevin steps/s on the host
        line["langevin_rk_kpm"] = cpu_langevin_baseline()
    except Exception as exc:
        line["langevin_rk_kpm"] = {"error": str(exc)[:200]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--replicas", type=int, default=256, help="replicas per launch of the device-resident step")
    ap.add_argument("--launches-per-step", type=int, default=50,
                    help="a step = this many passes of the fused M^T M kernel over the replica batch (a single 0.22 ms launch per "
                         "step made the timed window shorter than the power-management time scale)")
    ap.add_argument("--e2e-replicas", type=int, default=64, help="replicas per host-buffer (e2e) step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the CG / Langevin extra measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly one JSON line: NCCL's own messages (the image sets NCCL_DEBUG=VERSION, whose banner goes to
    # stdout) are sent to stderr, and the bare version banner is dropped; must happen before NCCL is first touched
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"

    import ctypes as C
    import torch
    import torch.distributed as dist
    import elphdynamics_b200 as E
    from elphdynamics_b200._lib import ptr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    all_cores = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    # work on a real (non-legacy) stream: CUDA events time it, and the engine can capture CUDA graphs on it
    torch.cuda.set_stream(torch.cuda.Stream())
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    em, rng = build_model()
    lib = em._lib
    stream = torch.cuda.current_stream()
    em.set_stream(stream.cuda_stream)
    n = em.Ndim
    Nsites, Ltau = em.Nsites, em.Ltau
    R = args.replicas
    hbm_peak, peak_src, _ = peaks()

    # ---------------- device-resident inputs: R replicas, each with its own expnV table and vector ------------
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    V = torch.randn(R, n, dtype=torch.float64, device="cuda", generator=g)
    Y = torch.empty_like(V)
    # expnV of the synthetic field in the engine layout [tau][site], perturbed per replica (independent chains)
    base = torch.from_numpy(np.ascontiguousarray(em.expnV.reshape(Nsites, Ltau).T)).reshape(-1).cuda()
    D = base.unsqueeze(0).repeat(R, 1) * (1.0 + 0.01 * torch.rand(R, n, dtype=torch.float64, device="cuda", generator=g))

    LPS = max(1, args.launches_per_step)

    def step_device():
        for _ in range(LPS):
            st = lib.elph_dev_mulMTM_replicas(em.handle, R, D.data_ptr(), n, V.data_ptr(), Y.data_ptr(), n)
            if st != 0:
                raise RuntimeError(lib.elph_last_error(em.handle))

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # the timed region is only a few milliseconds, shorter than nvidia-smi's sampling period: precede it with ~0.4 s
    # of the SAME launches (untimed) so that the clock / throttle record is taken under this very load
    t_soak = time.perf_counter()
    while time.perf_counter() - t_soak < 0.4:
        step_device()
        torch.cuda.synchronize()
    barrier()
    l0 = em.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = em.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * R * LPS * args.steps / (ms * 1e-3)
    kernel_us = ms * 1e3 / (args.steps * LPS)
    achieved = BYTES_PER_POINT * n * R / (kernel_us * 1e-6) / 1e9   # GB/s per GPU

    # ---------------- e2e: host buffers through the C ABI ------------------------------------------------------
    Re = args.e2e_replicas
    Vh = torch.randn(Re, n, dtype=torch.float64).pin_memory()
    Yh = torch.empty(Re, n, dtype=torch.float64).pin_memory()
    vp = C.cast(Vh.data_ptr(), C.POINTER(C.c_double))
    yp = C.cast(Yh.data_ptr(), C.POINTER(C.c_double))

    def step_e2e():
        st = lib.elph_mulMTM_batch(em.handle, Re, vp, yp)
        if st != 0:
            raise RuntimeError(lib.elph_last_error(em.handle))

    for _ in range(3):
        step_e2e()
    barrier()
    ksteps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(ksteps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * Re * ksteps / e2e_s

    extra = {}
    # ---- the second headline quantity on every GPU of the job: one independent Markov chain per rank (the reference's own
    # scale-out, ElPhDynamics.jl:90-95), Runge-Kutta Langevin steps through the C ABI with host noise, no data-path collective
    lang_all = None
    if not args.no_extra:
        try:
            from elphdynamics_b200 import workloads as _wl
            mL, rL = _wl.holstein("square", LSIDE, BETA, DTAU, mu=-1.0, seed=7000 + rank, eps=0.3)
            faL = E.FourierAccelerator(mL)
            E.update_Q_(faL, mL, 0.0, 10.0, 1.0)
            PL = E.SymmetricKPMPreconditioner(mL)
            dynL = E.RungeKuttaDynamics(mL, 1e-3)
            nstL = 20
            nzL = [dict(eta=rL.normal(size=n), g1=rL.normal(size=n), g2=rL.normal(size=n),
                        arnoldi1=rL.normal(size=2 * Nsites), arnoldi2=rL.normal(size=2 * Nsites)) for _ in range(nstL + 2)]
            for z in nzL:
                for key in ("eta", "g1", "g2"):
                    mL.pin_host(z[key])
            for z in nzL[nstL:]:
                E.evolve_(mL, dynL, faL, PL, **z)       # warm-up
            barrier()
            t0 = time.perf_counter()
            itsL = [E.evolve_(mL, dynL, faL, PL, **nzL[k]) for k in range(nstL)]
            dtL = time.perf_counter() - t0
            for z in nzL:
                for key in ("eta", "g1", "g2"):
                    mL.unpin_host(z[key])
            mL.close()
            if world > 1:
                t = torch.tensor([dtL], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtL = float(t.item())
            lang_all = {"chains": world, "steps_per_s_aggregate": world * nstL / dtL, "steps_per_s_per_chain": nstL / dtL,
                        "steps_timed_per_chain": nstL, "pcg_iters_last": itsL[-3:],
                        "note": "one 32x32xL200 chain per GPU, elph_langevin_step (Runge-Kutta, KPM-preconditioned, speculative "
                                "set-up) with page-locked host noise; time = max over ranks"}
        except Exception as exc:                          # an extra figure must not cost the bench line
            lang_all = {"error": str(exc)[:300]}
    if lang_all is not None:
        extra["langevin_rk_kpm_per_gpu_chains"] = lang_all
    if not args.no_extra and rank == 0:
        # single lattice, L2-resident: latency-bound regime of the real simulation
        v1 = V[0].contiguous()
        y1 = torch.empty_like(v1)
        for _ in range(20):
            lib.elph_dev_mulMTM(em.handle, v1.data_ptr(), y1.data_ptr())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            lib.elph_dev_mulMTM(em.handle, v1.data_ptr(), y1.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        us1 = e0.elapsed_time(e1) * 1e3 / 200
        extra["single_lattice"] = {"us_per_matvec": us1, "matvecs_per_s": 1e6 / us1,
                                   "algorithmic_GBps": BYTES_PER_POINT * n / us1 / 1e3, "note": "L2-resident, launch/latency bound"}
        # several right-hand sides on ONE field (the two spins of HMC, the n_v measurement vectors): expnV read once
        multi = {}
        for nrhs in (2, 10):
            Vm = V[:nrhs].contiguous()
            Ym = torch.empty_like(Vm)
            for _ in range(20):
                lib.elph_dev_mulMTM_replicas(em.handle, nrhs, None, 0, Vm.data_ptr(), Ym.data_ptr(), n)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(200):
                lib.elph_dev_mulMTM_replicas(em.handle, nrhs, None, 0, Vm.data_ptr(), Ym.data_ptr(), n)
            e1.record()
            torch.cuda.synchronize()
            usm = e0.elapsed_time(e1) * 1e3 / 200
            multi[str(nrhs)] = {"us_per_launch": usm, "matvecs_per_s": nrhs * 1e6 / usm}
        extra["multi_rhs_one_field"] = multi
        # CG and one Langevin RK step with injected noise (KPM-preconditioned), through the C ABI
        gvec = rng.normal(size=n)
        b = np.zeros(n)
        E.mulMT_(b, em, gvec)
        xs = np.zeros(n)
        E.ldiv_(xs, em, b)                    # warm-up: one-time arena allocation and kernel attributes of the persistent CG
        xs[:] = 0.0
        t0 = time.perf_counter()
        it, res, flag = E.ldiv_(xs, em, b)
        dt_cg = time.perf_counter() - t0
        extra["cg"] = {"iters": it, "residual": res, "flag": flag, "seconds": dt_cg, "iters_per_s": it / dt_cg,
                       "note": "elph_solve (ldiv!) with host b, x: single-reduction persistent kernel (one grid barrier per "
                               "iteration) + true-residual check + copies"}
        # measurement solves: n_v = 10 right-hand sides on one field in one call (host buffers through the C ABI)
        Bm = rng.normal(size=(10, n))
        Xm = np.zeros_like(Bm)
        E.ldiv_batch_(Xm, em, Bm)
        t0 = time.perf_counter()
        infos = E.ldiv_batch_(Xm, em, Bm)
        dt_b = time.perf_counter() - t0
        extra["solve_batch_10rhs"] = {"seconds": dt_b, "solves_per_s": 10 / dt_b, "iters": [i[0] for i in infos],
                                      "cg_iters_per_s": sum(i[0] for i in infos) / dt_b,
                                      "note": "elph_solve_batch: right-hand sides share one persistent cooperative CG launch "
                                              "(two at a time fit the machine at 32x32xL200)"}
        P = E.SymmetricKPMPreconditioner(em)
        kinfo = E.setup_(P, rng.normal(size=2 * Nsites))
        for _ in range(10):
            lib.elph_dev_kpm_apply(em.handle, v1.data_ptr(), y1.data_ptr())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            lib.elph_dev_kpm_apply(em.handle, v1.data_ptr(), y1.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        usk = e0.elapsed_time(e1) * 1e3 / 100
        extra["kpm_apply"] = {"us_per_apply": usk, "applies_per_s": 1e6 / usk, "total_order": int(kinfo.total_order),
                              "max_order": int(kinfo.max_order), "sweeps_per_apply": 2 * int(kinfo.total_order),
                              "note": "latency bound (sequential depth 2*max_order); 2 FFT kernels + cluster-split recurrences"}
        xs = np.zeros(n)
        t0 = time.perf_counter()
        itp, resp, flagp = E.ldiv_(xs, em, b, P)
        dt_p = time.perf_counter() - t0
        # device-side time of the solve alone (device pointers, no copies, no true-residual check)
        bdev = torch.from_numpy(np.ascontiguousarray(b.reshape(Nsites, Ltau).T)).reshape(-1).cuda()
        xdev = torch.zeros(n, dtype=torch.float64, device="cuda")
        itc, epsc = C.c_int64(), C.c_double()
        pcg = {}
        for label, fused in (("one_persistent_kernel", 1), ("launch_per_phase", 0)):
            lib.elph_set_tuning(em.handle, 17, fused)
            best = float("inf")
            for _ in range(3):
                xdev.zero_()
                torch.cuda.synchronize()
                l0 = em.launch_count()
                t0 = time.perf_counter()
                lib.elph_dev_cg_solve(em.handle, bdev.data_ptr(), xdev.data_ptr(), 1, 0.0, 0, C.byref(itc), C.byref(epsc))
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
                nl = em.launch_count() - l0
            pcg[label] = {"iters": int(itc.value), "us_per_iteration": best * 1e6 / max(1, itc.value), "launches_per_solve": int(nl)}
        lib.elph_set_tuning(em.handle, 17, 1)
        extra["pcg_kpm"] = {"iters": itp, "residual": resp, "flag": flagp, "seconds": dt_p, **pcg,
                            "note": "KPM-preconditioned CG: ldiv! with host vectors (seconds), and the solve alone on device vectors as "
                                    "one persistent cooperative kernel (pcg_fused.cu: FFT / Chebyshev chains / inverse FFT / product "
                                    "between grid barriers) against four launches per iteration"}
        fa = E.FourierAccelerator(em)
        E.update_Q_(fa, em, 0.0, 10.0, 1.0)
        dyn = E.RungeKuttaDynamics(em, 1e-3)
        nsteps = 5
        its = []
        noise = [dict(eta=rng.normal(size=n), g1=rng.normal(size=n), g2=rng.normal(size=n),
                      arnoldi1=rng.normal(size=2 * Nsites), arnoldi2=rng.normal(size=2 * Nsites)) for _ in range(nsteps + 1)]
        for nz in noise:                             # the driver's preallocated noise vectors, page-locked once
            for key in ("eta", "g1", "g2"):
                em.pin_host(nz[key])
        # the timed window is short (5 steps = 16 ms): three repetitions of the SAME five steps (field reset to the start, warm-up
        # step, then the timed steps), best and all reported -- one host hiccup must not decide the figure
        x_start = em.x.copy()
        reps_l = []
        for rep in range(3):
            em.x = x_start
            E.update_model_(em)
            E.evolve_(em, dyn, fa, P, **noise[nsteps])   # warm-up
            t0 = time.perf_counter()
            its = [E.evolve_(em, dyn, fa, P, **noise[k]) for k in range(nsteps)]   # noise drawn outside the timed region
            reps_l.append(nsteps / (time.perf_counter() - t0))
        for nz in noise:
            for key in ("eta", "g1", "g2"):
                em.unpin_host(nz[key])
        extra["langevin_rk_kpm"] = {"steps_per_s": max(reps_l), "steps_per_s_repetitions": reps_l, "pcg_iters_second_solve": its,
                                    "note": "elph_langevin_step through the C ABI with host noise buffers (page-locked with "
                                            "elph_host_register); 2 KPM set-ups (speculative: the Arnoldi bounds are computed beside "
                                            "the solve) + 2 KPM-PCG solves + forces + Fourier acceleration; best of three repetitions "
                                            "of the same five steps"}

        from elphdynamics_b200 import workloads
        # ---- independent Markov chains on ONE GPU (the reference's own scale-out, ElPhDynamics.jl:90-95: one process per
        # chain id): a single chain is latency bound and leaves most of the device idle, so K chains, each with its own
        # handle, stream and host thread (ctypes releases the GIL inside the C ABI), advance concurrently
        try:
            import threading as _th
            nsm = torch.cuda.get_device_properties(local).multi_processor_count
            chain_results = {}
            # (chains, one-kernel solve?, CTAs per solve): many chains fill the GPU with the launch-per-phase solve; with the
            # persistent one-kernel solve every chain takes 1/K of the SMs (tuning key 20) and keeps a much shorter step
            for K, fused, grid in ((8, 0, 0), (4, 1, (nsm // 4) & ~1)):
                nst = 6
                chains = []
                for c in range(K):
                    mc, rc = workloads.holstein("square", LSIDE, BETA, DTAU, mu=-1.0, seed=4321 + c, eps=0.3)
                    fc = E.FourierAccelerator(mc)
                    E.update_Q_(fc, mc, 0.0, 10.0, 1.0)
                    Pc = E.SymmetricKPMPreconditioner(mc)
                    mc._call("elph_set_tuning", 17, fused)
                    mc._call("elph_set_tuning", 20, grid)
                    dc = E.RungeKuttaDynamics(mc, 1e-3)
                    nz = [dict(eta=rc.normal(size=n), g1=rc.normal(size=n), g2=rc.normal(size=n),
                               arnoldi1=rc.normal(size=2 * Nsites), arnoldi2=rc.normal(size=2 * Nsites)) for _ in range(nst + 1)]
                    for z in nz:
                        for key in ("eta", "g1", "g2"):
                            mc.pin_host(z[key])
                    chains.append((mc, fc, Pc, dc, nz))
                its_c = [0] * K

                def run_chain(c, first, steps):
                    mc, fc, Pc, dc, nz = chains[c]
                    for k in range(first, first + steps):   # fresh noise every step, as in a real chain
                        its_c[c] = E.evolve_(mc, dc, fc, Pc, **nz[k])

                for first, steps in ((0, 1), (1, nst)):            # warm-up step, then the timed steps
                    ths = [_th.Thread(target=run_chain, args=(c, first, steps)) for c in range(K)]
                    t0 = time.perf_counter()
                    for t in ths:
                        t.start()
                    for t in ths:
                        t.join()
                    dt_c = time.perf_counter() - t0
                chain_results[f"{K}_chains_" + ("one_kernel_solve" if fused else "launch_per_phase_solve")] = {
                    "chains": K, "steps_per_s_aggregate": K * nst / dt_c, "steps_per_s_per_chain": nst / dt_c,
                    "ctas_per_solve": grid or None, "pcg_iters_last": list(its_c)}
                for mc, fc, Pc, dc, nz in chains:
                    for z in nz:
                        for key in ("eta", "g1", "g2"):
                            mc.unpin_host(z[key])
                    mc.close()
            best = max(chain_results.values(), key=lambda r: r["steps_per_s_aggregate"])
            extra["langevin_rk_kpm_chains"] = {"chains": best["chains"], "steps_per_s_aggregate": best["steps_per_s_aggregate"],
                                               "steps_per_s_per_chain": best["steps_per_s_per_chain"], "variants": chain_results,
                                               "note": "K independent 32x32xL200 chains on one GPU, one handle + stream + host thread each "
                                                       "(elph_langevin_step, page-locked host noise)"}
        except Exception as exc:   # an extra must not cost the bench line
            extra["langevin_rk_kpm_chains"] = {"error": str(exc)[:200]}

        # ---- measurement side: the four convolutions of setup!(estimator, n1, n2) on the device ----
        from elphdynamics_b200 import greens as eg
        Gr = eg.EstimateGreensFunction(em, 4)
        Gr.R[:] = rng.normal(size=Gr.R.shape)
        Gr.MinvR[:] = rng.normal(size=Gr.R.shape)     # stand-ins for the solves: the cost of a pair does not depend on them
        em._call("elph_greens_load", Gr.nv, ptr(Gr.R), ptr(Gr.MinvR))
        eg.setup_pair_(Gr, 0, 1)
        t0 = time.perf_counter()
        for (i1, i2) in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)):
            eg.setup_pair_(Gr, i1, i2)
        dt_g = (time.perf_counter() - t0) / 6
        extra["greens_setup_pair"] = {"ms_per_pair": dt_g * 1e3, "pairs_per_s": 1.0 / dt_g,
                                      "note": "setup!(estimator, n1, n2): 4 convolutions = 12 transforms over (2 Ltau, L1, L2) on "
                                              "the device, 4 x 6.5 MB back to the host (elph_greens_setup)"}
        del Gr

        # ---- configuration C: SSH 32x32xL200 (per-(tau,bond) cosh/sinh tables, 48 B/pt algorithmic) ----
        from elphdynamics_b200 import hmc as ehmc
        mC, rC = workloads.config("C")
        mC.set_stream(stream.cuda_stream)
        nC = mC.Ndim
        vC = torch.randn(nC, dtype=torch.float64, device="cuda")
        yC = torch.empty_like(vC)
        for _ in range(10):
            lib.elph_dev_mulMTM(mC.handle, vC.data_ptr(), yC.data_ptr())
        torch.cuda.synchronize()
        usC = float("inf")
        for _ in range(3):     # launch-bound (one 6 us kernel per call): best of three runs of 200 launches
            e0.record()
            for _ in range(200):
                lib.elph_dev_mulMTM(mC.handle, vC.data_ptr(), yC.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            usC = min(usC, e0.elapsed_time(e1) * 5.0)
        bC = rC.normal(size=nC)
        xC = np.zeros(nC)
        t0 = time.perf_counter()
        itC, resC, flC = E.ldiv_(xC, mC, bC)
        dtC = time.perf_counter() - t0
        extra["ssh_square_32x32_L200"] = {"us_per_matvec": usC, "matvecs_per_s": 1e6 / usC,
                                          "algorithmic_GBps": 48.0 * nC / usC / 1e3,
                                          "cg": {"iters": int(itC), "residual": float(resC), "flag": int(flC), "seconds": dtC},
                                          "note": "single lattice (L2-resident tables), best of 3 x 200 launches; 48 B/pt = v + cosh + sinh tables + y"}
        # HBM regime: independent replicas, each with its own phonon field -> own (cosh, sinh) table (6.6 MB) and vector
        RC = 128
        LC, NC, NphC = mC.Ltau, mC.Nsites, mC.Nph
        tab_stride = 4 * LC * NC
        XC = 0.3 * torch.randn(RC, NphC * LC, dtype=torch.float64, device="cuda")
        TC = torch.empty(RC * tab_stride, dtype=torch.float64, device="cuda")
        VC = torch.randn(RC, nC, dtype=torch.float64, device="cuda")
        YC = torch.empty_like(VC)
        e0.record()
        mC._call("elph_dev_ssh_replica_tables", RC, XC.data_ptr(), NphC * LC, TC.data_ptr(), tab_stride)
        e1.record()
        torch.cuda.synchronize()
        us_tab = e0.elapsed_time(e1) * 1e3
        for _ in range(5):
            lib.elph_dev_mulMTM_replicas_ssh(mC.handle, RC, TC.data_ptr(), tab_stride, VC.data_ptr(), YC.data_ptr(), nC)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            lib.elph_dev_mulMTM_replicas_ssh(mC.handle, RC, TC.data_ptr(), tab_stride, VC.data_ptr(), YC.data_ptr(), nC)
        e1.record()
        torch.cuda.synchronize()
        us_rep = e0.elapsed_time(e1) * 1e3 / 20
        extra["ssh_square_32x32_L200"]["replicas"] = {
            "replicas_per_launch": RC, "us_per_launch": us_rep, "matvecs_per_s": RC * 1e6 / us_rep,
            "algorithmic_bytes_per_launch": 48.0 * nC * RC, "achieved_GBps": 48.0 * nC * RC / us_rep / 1e3,
            "frac_of_hbm_peak": 48.0 * nC * RC / us_rep / 1e3 / hbm_peak, "peak": hbm_peak, "peak_source": peak_src,
            "tables_us": us_tab,
            "note": "elph_dev_mulMTM_replicas_ssh: every replica streams v + (cosh, sinh) of both bond directions + y = 48 B per "
                    "lattice point (805 MB per launch vs 126 MB L2); tables_us = update_model! of all replicas in one launch "
                    "(elph_dev_ssh_replica_tables)"}
        del XC, TC, VC, YC
        # the second Langevin configuration of BASELINE.json (examples/ssh_langevin_square.toml scaled to 32x32, L = 200):
        # Runge-Kutta steps with Fourier acceleration (mass 0.1) and the KPM-preconditioned solve as one persistent kernel
        try:
            PC = E.SymmetricKPMPreconditioner(mC)
            faC = E.FourierAccelerator(mC)
            E.update_Q_(faC, mC, 0.0, 10.0, 0.1)
            dynC = E.RungeKuttaDynamics(mC, 1e-3)
            nzC = [dict(eta=rC.normal(size=mC.Ndof), g1=rC.normal(size=nC), g2=rC.normal(size=nC),
                        arnoldi1=rC.normal(size=2 * NC), arnoldi2=rC.normal(size=2 * NC)) for _ in range(6)]
            for z in nzC:
                for key in ("eta", "g1", "g2"):
                    mC.pin_host(z[key])
            E.evolve_(mC, dynC, faC, PC, **nzC[0])
            t0 = time.perf_counter()
            itsC = [int(E.evolve_(mC, dynC, faC, PC, **z)) for z in nzC[1:]]
            dtC = (time.perf_counter() - t0) / 5
            for z in nzC:
                for key in ("eta", "g1", "g2"):
                    mC.unpin_host(z[key])
            extra["ssh_square_32x32_L200"]["langevin_rk_kpm"] = {
                "steps_per_s": 1.0 / dtC, "pcg_iters_second_solve": itsC,
                "note": "elph_langevin_step, SSH model: 2 KPM set-ups + 2 KPM-PCG solves (one persistent kernel each, per-slice "
                        "(cosh, sinh) tables resident in shared memory) + forces + Fourier acceleration"}
        except Exception as exc:
            extra["ssh_square_32x32_L200"]["langevin_rk_kpm"] = {"error": str(exc)[:200]}
        mC.close()

        # ---- configuration D: HMC trajectory on the honeycomb lattice L=32 (N=2048, Ltau=20), Nt=10 leapfrog steps ----
        mD, rD = workloads.config("D_honeycomb")
        mD.set_stream(stream.cuda_stream)
        faD = E.FourierAccelerator(mD)
        E.update_M_(faD, mD, 0.0, 10.0, 1.0, 0.0)
        hD = ehmc.HybridMonteCarlo(mD, 0.01, 0.1, 0.0, 10)
        draws = [dict(R_v=rD.normal(size=mD.Ndof), R_plus=rD.normal(size=mD.Ndim), R_minus=rD.normal(size=mD.Ndim),
                      uniform=float(rD.uniform())) for _ in range(6)]
        ehmc.update_(mD, hD, faD, None, **draws[0])
        t0 = time.perf_counter()
        itsD, accD = [], []
        for d in draws[1:]:
            a, it = ehmc.update_(mD, hD, faD, None, **d)
            itsD.append(float(it)); accD.append(bool(a))
        dtD = (time.perf_counter() - t0) / (len(draws) - 1)
        extra["hmc_honeycomb_L32"] = {"trajectories_per_s": 1.0 / dtD, "leapfrog_steps": hD.Nt, "inner_Sb_steps": hD.Nb,
                                      "solves_per_trajectory": 2 * (hD.Nt + 2), "cg_iters": itsD, "accepted": accD,
                                      "note": "elph_hmc_update: whole trajectory on the device, noise injected from the host"}
        # the lattice's kernels in isolation: one M^T M, one CG solve (persistent kernel), and the HBM regime with independent
        # replicas (own expnV table and vector each: 0.98 MB per replica, 1024 replicas = 1 GB per launch)
        nD = mD.Ndim
        vD = torch.randn(nD, dtype=torch.float64, device="cuda")
        yD = torch.empty_like(vD)
        for _ in range(10):
            lib.elph_dev_mulMTM(mD.handle, vD.data_ptr(), yD.data_ptr())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            lib.elph_dev_mulMTM(mD.handle, vD.data_ptr(), yD.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        usD = e0.elapsed_time(e1) * 1e3 / 200
        bD = torch.randn(nD, dtype=torch.float64, device="cuda")
        xD = torch.zeros_like(bD)
        itD, epsD = C.c_int64(), C.c_double()
        bestD = float("inf")
        for _ in range(3):
            xD.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lib.elph_dev_cg_solve(mD.handle, bD.data_ptr(), xD.data_ptr(), 0, 0.0, 0, C.byref(itD), C.byref(epsD))
            torch.cuda.synchronize()
            bestD = min(bestD, time.perf_counter() - t0)
        RD = 1024
        DD = torch.from_numpy(np.ascontiguousarray(mD.expnV.reshape(mD.Nsites, mD.Ltau).T)).reshape(-1).cuda()
        DD = DD.unsqueeze(0).repeat(RD, 1) * (1.0 + 0.01 * torch.rand(RD, nD, dtype=torch.float64, device="cuda"))
        VD = torch.randn(RD, nD, dtype=torch.float64, device="cuda")
        YD = torch.empty_like(VD)
        for _ in range(3):
            lib.elph_dev_mulMTM_replicas(mD.handle, RD, DD.data_ptr(), nD, VD.data_ptr(), YD.data_ptr(), nD)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            lib.elph_dev_mulMTM_replicas(mD.handle, RD, DD.data_ptr(), nD, VD.data_ptr(), YD.data_ptr(), nD)
        e1.record()
        torch.cuda.synchronize()
        usRD = e0.elapsed_time(e1) * 1e3 / 20
        extra["hmc_honeycomb_L32"].update({
            "us_per_matvec": usD, "cg": {"iters": int(itD.value), "us_per_iteration": bestD * 1e6 / max(1, itD.value)},
            "replicas": {"replicas_per_launch": RD, "us_per_launch": usRD, "matvecs_per_s": RD * 1e6 / usRD,
                         "achieved_GBps": BYTES_PER_POINT * nD * RD / usRD / 1e3,
                         "frac_of_hbm_peak": BYTES_PER_POINT * nD * RD / usRD / 1e3 / hbm_peak,
                         "note": "register-tile kernels for the honeycomb lattice (cell pair / lane rotation / row pair); 24 B per "
                                 "lattice point, 1.0 GB per launch"}})
        del DD, VD, YD
        mD.close()

    # ---------------- tau-sharded single lattice (config E: Holstein 64x64, L=400), strong scaling over ranks -----
    sharded = None
    if not args.no_extra:
        from elphdynamics_b200.sharded import CudaSlabBackend, RingComm, ShardedOperator, slab_bounds
        LE, LtauE = 64, 400
        tau0, lloc = slab_bounds(LtauE, world, rank)
        latE = E.Lattice(E.UnitCell(2, 1), LE)
        mE = E.HolsteinModel(latE, lloc * DTAU, DTAU, tol=1e-5, maxiter=10000)
        mE.assign_omega(1.0); mE.assign_lambda(1.0); mE.assign_mu(-1.0)
        mE.assign_t(1.0, 0, 0, (1, 0, 0)); mE.assign_t(1.0, 0, 0, (0, 1, 0))
        mE.initialize_model_()
        # the same global synthetic field whatever the world size: generated globally, sliced per rank (host layout site-major)
        rs = np.random.default_rng(99)
        xgE = rs.integers(-1, 2, size=(mE.Nsites, 1)) + 0.7 * rs.normal(size=(mE.Nsites, 1)) + 0.3 * rs.normal(size=(mE.Nsites, LtauE))
        bgE = rs.normal(size=(LtauE, mE.Nsites))
        mE.x = np.ascontiguousarray(xgE[:, tau0:tau0 + lloc]).reshape(-1)
        beE = CudaSlabBackend(mE, tau0, LtauE)
        opE = ShardedOperator(beE, RingComm(rank, world))
        opE.update_model()
        # products first with the torch.distributed halos (NCCL send/recv), then with the halos pushed through peer memory inside
        # one kernel (elph_dev_shard_halo): no NCCL call, no host synchronisation per product
        vE, yE = beE.empty(), beE.empty()
        vE[1:lloc + 1].normal_()

        def time_products(nrep=50):
            for _ in range(5):
                opE.mulMTM(yE, vE)
            barrier()
            e0.record()
            for _ in range(nrep):
                opE.mulMTM(yE, vE)
            e1.record()
            barrier()
            ms_ = e0.elapsed_time(e1)
            if world > 1:
                t_ = torch.tensor([ms_], dtype=torch.float64, device="cuda")
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
                ms_ = float(t_.item())
            return ms_ * 1e3 / nrep

        us_nccl = time_products()
        p2p_open = opE.enable_p2p()
        usE = time_products() if p2p_open else us_nccl
        sharded = {"workload": "holstein_square_64x64_L400, one lattice tau-sharded over the ranks (strong scaling)",
                   "us_per_matvec": usE, "matvecs_per_s": 1e6 / usE, "slab_slices_per_gpu": lloc,
                   "us_per_matvec_nccl_halo": us_nccl,
                   "algorithmic_GBps_per_gpu": BYTES_PER_POINT * mE.Nsites * lloc / usE / 1e3,
                   "collective": ("1 halo slice each way per product pushed through NVLink peer memory inside one kernel "
                                  "(elph_dev_shard_halo)" if p2p_open else "1 halo slice each way per product (NCCL send/recv)") +
                                 ", antiperiodic sign on global slice 0"}
        # CG on the sharded lattice: the pipelined persistent kernel (halo pushes + all-reduce over NVLink inside the kernel,
        # csrc/cg_pipe.cu) where every slab is co-resident, else the single-GPU engine (world = 1) or the NCCL-between-launches loop
        bE, xE = beE.empty(), beE.empty()
        bE[1:lloc + 1] = torch.from_numpy(bgE[tau0:tau0 + lloc]).cuda()
        cgE = {}
        if getattr(beE, "_p2p_ready", False):
            opE.solve(xE, bE)
            barrier()
            t0 = time.perf_counter()
            itE, epsE = opE.solve(xE, bE)
            torch.cuda.synchronize()
            dtE = time.perf_counter() - t0
            kv = C.c_int32()
            lib.elph_get_tuning(mE.handle, 100, C.byref(kv))
            cgE = {"path": "persistent kernel per GPU: halo pushes + scalar all-reduce over NVLink inside the kernel",
                   "kernel": f"cgpipe variant*100 + CTAs per slice*10 + warps = {kv.value}" if kv.value else "cg_p2p (single reduction)"}
        elif world == 1:
            from elphdynamics_b200 import workloads
            mF, _ = workloads.holstein("square", LE, LtauE * DTAU, DTAU, seed=5)
            mF.x = xgE.reshape(-1)
            E.update_model_(mF)
            mF.set_stream(torch.cuda.current_stream().cuda_stream)
            bF = torch.from_numpy(np.ascontiguousarray(bgE).reshape(-1)).cuda()     # engine layout [tau][site], as the slabs
            xF = torch.zeros_like(bF)
            itc, epc = C.c_int64(), C.c_double()
            for _ in range(2):
                xF.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                mF._lib.elph_dev_cg_solve(mF.handle, bF.data_ptr(), xF.data_ptr(), 0, 0.0, 0, C.byref(itc), C.byref(epc))
                torch.cuda.synchronize()
                dtE = time.perf_counter() - t0
            itE, epsE = itc.value, epc.value
            cgE = {"path": "single-GPU engine (CUDA-graph CG: 400 slices of 64x64 are not co-resident on one GPU)"}
            mF.close()
        else:
            barrier()
            t0 = time.perf_counter()
            itE, epsE = opE.solve_cg(xE, bE, maxiter=40)
            torch.cuda.synchronize()
            dtE = time.perf_counter() - t0
            cgE = {"path": "launch-per-operation CG, halos through peer memory, scalar all-reduces through NCCL (bounded to 40 iterations)"}
        if world > 1:
            t = torch.tensor([dtE], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtE = float(t.item())
        cgE.update({"iters": int(itE), "eps": float(epsE), "seconds": dtE, "us_per_iter": dtE / max(int(itE), 1) * 1e6,
                    "algorithmic_GBps_aggregate": 96.0 * mE.Nsites * LtauE * int(itE) / dtE / 1e9})
        # strong-scaling efficiency of this run against the committed 1-GPU figures of the same lattice (this process only knows
        # its own N; the driver's SCALE record holds all four runs)
        ref_p = ROOT / "profiles" / "tau_sharded_reference.json"
        if ref_p.exists():
            try:
                ref = json.loads(ref_p.read_text())
                sharded["efficiency"] = {
                    "cg_strong": ref["cg_us_per_iter_1gpu"] / (world * cgE["us_per_iter"]),
                    "matvec_strong": ref["us_per_matvec_1gpu"] / (world * usE),
                    "reference": {"cg_us_per_iter_1gpu": ref["cg_us_per_iter_1gpu"], "us_per_matvec_1gpu": ref["us_per_matvec_1gpu"],
                                  "source": ref.get("source", "profiles/tau_sharded_reference.json")}}
            except Exception:
                pass
        sharded["cg"] = cgE
        # KPM-preconditioned CG on the same sharded lattice: omega-sharded application of the preconditioner, the three transposes
        # pulled through peer memory inside the stage kernels (csrc/kpm_shard.cu), products with the halo inside the kernel
        try:
            from elphdynamics_b200.sharded import ShardedKPM
            auxE = E.HolsteinModel(latE, LtauE * DTAU, DTAU, tol=1e-5, maxiter=10000)
            auxE.assign_omega(1.0); auxE.assign_lambda(1.0); auxE.assign_mu(-1.0)
            auxE.assign_t(1.0, 0, 0, (1, 0, 0)); auxE.assign_t(1.0, 0, 0, (0, 1, 0))
            auxE.initialize_model_()
            beE.kpm_init(auxE)
            PE = ShardedKPM(opE, mE.Nsites, LtauE)
            fusedE = PE.enable_fused(tau0)
            noiseE = np.random.default_rng(98).normal(size=2 * mE.Nsites)
            PE.setup(noiseE)
            zE = beE.empty()
            for _ in range(3):
                PE.ldiv(zE, bE)
            barrier()
            e0.record()
            for _ in range(10):
                PE.ldiv(zE, bE)
            e1.record()
            barrier()
            us_apply = e0.elapsed_time(e1) * 1e3 / 10
            xE.zero_()
            opE.solve_pcg(xE, bE, PE)
            barrier()
            xE.zero_()
            t0 = time.perf_counter()
            itP, epsP = opE.solve_pcg(xE, bE, PE)
            torch.cuda.synchronize()
            dtP = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dtP, us_apply], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtP, us_apply = float(t[0].item()), float(t[1].item())
            if fusedE:
                beE.kpm_shard_check()
            sharded["pcg_kpm"] = {
                "iters": int(itP), "eps": float(epsP), "seconds": dtP, "us_per_iter": dtP / max(int(itP), 1) * 1e6,
                "us_per_kpm_apply": us_apply, "frequencies_per_gpu": len(PE.my_w), "max_order": int(beE.kpm_orders().max()),
                "transposes": "pulled through peer memory inside the FFT / gather kernels, 4 barrier kernels per application "
                              "(csrc/kpm_shard.cu)" if fusedE else "4 NCCL all-to-alls per application",
                "note": "ShardedOperator.solve_pcg: host-driven loop (product with the halo inside the kernel, preconditioner "
                        "application queued before |r|^2 is read back: 2 scalar round trips per iteration); the application is bounded by the Chebyshev chain of the "
                        "lowest frequency (2 x max_order dependent sweeps of one 64x64 slice), which omega-sharding does not shorten"}
            auxE.close()
        except Exception as exc:              # an extra figure must not cost the bench line
            sharded["pcg_kpm"] = {"error": str(exc)[:300]}
        mE.close()

    if rank == 0:
        traffic, traffic_src = None, None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                tj = json.loads(tp.read_text())
                # captured at tj["replicas"] replicas per launch; the kernel streams, so traffic is linear in replicas
                traffic = tj["mtm_replicas_dram_bytes_per_launch"] * R / tj.get("replicas", 64)
                traffic_src = (f"ncu --set full capture of this launch shape at {tj.get('replicas', 64)} replicas "
                               f"({tj.get('source', 'profiles/traffic.json')}), dram__bytes_read.sum + dram__bytes_write.sum; "
                               "not re-measured by this run (a profiler cannot run inside the timed region)")
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "holstein_square_32x32_L200", "replicas_per_gpu": R, "launches_per_step": LPS,
                           "matvecs_per_step": world * R * LPS,
                           "l2_policy": f"inputs larger than L2: {3 * R * n * 8 / 1e6:.0f} MB per launch vs 126 MB L2",
                           "parallelism": f"replicas x{world}" if world > 1 else "single GPU"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                             "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "mtm_square_kernel<1,16,0,256> (fused M^T M, register/shuffle, TMA-staged)",
                             "algorithmic_bytes_per_launch": BYTES_PER_POINT * n * R, "kernel_us": kernel_us},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": Re * n * 8, "d2h_bytes_per_step": Re * n * 8,
                        "api": "elph_mulMTM_batch (host pointers, pinned)", "host_numa_binding": numa},
                "gpu_launches": int(launches), "clocks": clocks}
        line.update(extra)
        if sharded is not None:
            line["tau_sharded"] = sharded
        if not args.no_cpu:
            os.sched_setaffinity(0, all_cores)     # the CPU baseline uses every host core, not only the GPU's NUMA node
            cb, _ = cpu_baseline()
            try:
                cb["langevin_rk_kpm"] = cpu_langevin_baseline()
            except Exception as exc:          # the extra CPU figure must not cost the bench line
                cb["langevin_rk_kpm"] = {"error": str(exc)[:200]}
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    em.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
