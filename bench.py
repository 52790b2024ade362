#!/usr/bin/env python
"""bench.py — batched factor+solve systems/s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5|cfg2|cfg3|cfg4]
    python bench.py --impl reference ...      # the reference's CPU path on the host cores

A step = one pass of the hot path over one batch of synthetic systems:
  cfg5 (default, the configuration the 1/2/4/8-GPU metric is quoted on): SAX-style
        photonic S-matrix systems, complex128, n_col=128, n_lhs=65536 PER GPU (weak
        scaling, systems sharded with no data-path collective), solve_with_symbol, n_rhs=1
  cfg2: same generator at n_col=64, n_lhs=4096 (working set < L2: L2 flushed between steps)
  cfg4: MNA/BTF float64 n_col=20000 n_lhs=128: refactor + solve + transpose solve per step
  cfg3: Poisson 256^2 float64 n_lhs=64: factor once, then solve_with_numeric n_rhs=256 per step
`value` is timed with CUDA events on the launching stream with inputs resident in HBM, calling
the C-ABI (include/klujax_cuda.h) directly; `e2e` goes through the public API
(klujax_b200.solve_with_symbol ...) from pinned HOST buffers with the copies inside the timed
region, next to the box's own pinned-copy ceiling for the same bytes (`e2e.pcie`).  Every number
carries a parity gate: a sample of the systems of the timed batch is solved by the CPU oracle
(`parity`: max relative error of x, max relative residual, systems checked).  The default run also
measures the other named configs on rank 0 (`extra_workloads`).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from klujax_b200 import synth  # noqa: E402

METRIC = "batched factor+solve systems/s"
UNIT = "systems/s"


# ----------------------------------------------------------------------------- workloads
class Workload:
    def __init__(self, name: str, n_lhs: int | None, rank: int = 0):
        self.name = name
        if name in ("cfg5", "cfg2"):
            self.n = 128 if name == "cfg5" else 64
            self.pat = synth.SaxPattern(self.n, seed=2 if name == "cfg5" else 1)
            self.L = n_lhs or (65536 if name == "cfg5" else 4096)
            self.r = 1
            self.dtype = np.complex128
            self.Ai, self.Aj = self.pat.Ai, self.pat.Aj
            self.desc = (f"SAX-style photonic S-matrix systems A=I-C.S, complex128, n_col={self.n}, "
                         f"n_nz={self.pat.nnz}, n_lhs={self.L} per GPU, n_rhs=1, solve_with_symbol")
            self.rank = rank
        elif name == "cfg4":
            self.pat = synth.MnaPattern()
            self.n, self.L, self.r, self.dtype = self.pat.n, n_lhs or 128, 1, np.float64
            self.Ai, self.Aj = self.pat.Ai, self.pat.Aj
            self.desc = (f"MNA-like BTF circuit matrices, float64, n_col={self.n}, n_nz={self.pat.nnz}, "
                         f"n_lhs={self.L} per GPU: refactor + solve + transpose solve per step")
        elif name == "cfg3":
            self.g = 256
            self.n, self.L, self.r, self.dtype = self.g * self.g, n_lhs or 64, 256, np.float64
            self.Ai, self.Aj, self._Ax = synth.poisson2d(self.g, self.L)
            self.desc = (f"2-D 5-point Poisson {self.g}^2, float64, n_lhs={self.L} per GPU, factor once, "
                         f"solve_with_numeric n_rhs={self.r} per step")
        else:
            raise SystemExit(f"unknown workload {name}")
        self.nz = int(self.Ai.size)
        self.es = 16 if self.dtype is np.complex128 else 8

    def values(self, count: int, seed: int = 0) -> np.ndarray:
        if self.name in ("cfg5", "cfg2"):
            return self.pat.values_fast(count, seed=12345 + seed)
        if self.name == "cfg4":
            return self.pat.values(seed, count)
        return self._Ax[:count]

    def rhs(self, count: int, seed: int = 0) -> np.ndarray:
        rng = np.random.default_rng(777 + seed)
        b = rng.standard_normal((count, self.n, self.r))
        if self.dtype is np.complex128:
            b = b + 1j * rng.standard_normal((count, self.n, self.r))
        return b.astype(self.dtype)

    def algorithmic_bytes_per_system(self, nnz_lu: int | None = None) -> float:
        """compulsory traffic only (indices/schedule are shared by the batch)."""
        s, z, n, r = self.es, self.nz, self.n, self.r
        if self.name in ("cfg5", "cfg2"):
            return s * (z + 2 * n * r)                       # E1 fused factor+solve
        if self.name == "cfg4":                              # E2 + 2 x E3
            lu = nnz_lu or 0
            return (s * (z + lu) + 8 * n) + 2 * (s * lu + 8 * n + 2 * s * n * r)
        lu = nnz_lu or 0
        return s * lu + 8 * n + 2 * s * n * r                # E3


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc, self.lines, self.gpu = None, [], gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- parity gate
def parity_sample(w: "Workload", Ax_np, b_np, x_np, n_check: int, xt_np=None) -> dict:
    """Solves a sample of the timed batch with the CPU oracle (per-system klu_factor + klu_solve,
    klujax.cpp:431-451) and returns the gate BASELINE.json asks for: x within 1e-10, residual
    within 1e-12.  Test infrastructure as the checker, never the thing measured."""
    from oracle import oracle as O
    L = Ax_np.shape[0]
    idx = np.unique(np.linspace(0, L - 1, min(n_check, L)).astype(int))
    sym = O.analyze(w.Ai, w.Aj, w.n)
    ref = O.solve_with_symbol(w.Ai, w.Aj, Ax_np[idx], b_np[idx], sym)
    err = float(np.abs(x_np[idx] - ref).max() / np.abs(ref).max())
    if xt_np is not None:
        reft = O.solve_with_symbol(w.Ai, w.Aj, Ax_np[idx], b_np[idx], sym, transpose=True)
        err = max(err, float(np.abs(xt_np[idx] - reft).max() / np.abs(reft).max()))
    O.free_symbolic(sym)
    res = 0.0
    for q, m in enumerate(idx[: min(len(idx), 64)]):
        Axx = np.zeros_like(b_np[m])
        np.add.at(Axx, w.Ai, Ax_np[m][:, None] * x_np[m][w.Aj])
        res = max(res, float(np.linalg.norm(Axx - b_np[m]) / np.linalg.norm(b_np[m])))
    return {"max_relerr": err, "max_residual": res, "n_checked": int(len(idx)),
            "x_tol": 1e-10, "residual_tol": 1e-12, "ok": bool(err <= 1e-10 and res <= 1e-12),
            "checker": "oracle/ (restated KLU, per-system klu_factor + klu_solve) on a sample of the timed batch"}


def lib_fingerprint() -> str:
    """Fingerprint of the CUDA library's SOURCES (csrc/ + include/): what an ncu capture in profiles/ was
    taken with.  (The binary itself is rebuilt by build() wherever the sources are newer.)"""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "klujax_b200", "csrc")
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".cc", ".h")))
    files.append(os.path.join(ROOT, "include", "klujax_cuda.h"))
    for path in files:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def bind_to_gpu_numa(local_rank: int) -> str:
    """Pins this rank's threads (and with them its pinned allocations, first touch) to the NUMA
    node of its GPU, so that eight ranks do not all stage through one socket."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:  # virtualised PCI topology: say what the box offers instead of guessing
            try:
                online = open("/sys/devices/system/node/online").read().strip()
            except OSError:
                online = "?"
            return f"GPU's numa node not exposed (nodes online: {online}; {len(os.sched_getaffinity(0))} cpus allowed): nothing to bind"
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return f"node {node} ({len(ids)} cpus)"
        return f"node {node} (no allowed cpu there)"
    except Exception as e:  # best effort
        return f"not bound ({type(e).__name__})"


# ----------------------------------------------------------------------------- CPU arm
def cpu_solve_with_symbol(w: Workload, Ax, b, threads: int):
    """The reference's serial loop (klujax.cpp:431-451), restated in oracle/, spread over
    `threads` host threads by batch range (ctypes releases the GIL)."""
    from oracle import oracle as O
    sym = O.analyze(w.Ai, w.Aj, w.n)
    L = Ax.shape[0]
    x = np.empty_like(b)
    bounds = np.linspace(0, L, threads + 1).astype(int)
    t0 = time.perf_counter()
    if threads == 1:
        O.solve_with_symbol(w.Ai, w.Aj, Ax, b, sym, out=x)
    else:
        ths = [threading.Thread(target=O.solve_with_symbol, args=(w.Ai, w.Aj, Ax, b, sym),
                                kwargs=dict(lo=int(bounds[i]), hi=int(bounds[i + 1]), out=x))
               for i in range(threads)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    dt = time.perf_counter() - t0
    O.free_symbolic(sym)
    return dt, x


def cpu_step(w: Workload, Ax, b, threads: int, state: dict):
    """One CPU 'step' of the workload's path on a sample of systems; returns seconds."""
    from oracle import oracle as O
    if w.name in ("cfg5", "cfg2"):
        return cpu_solve_with_symbol(w, Ax, b, threads)[0]
    if "sym" not in state:
        state["sym"] = O.analyze(w.Ai, w.Aj, w.n)
        state["num"] = O.factor(w.Ai, w.Aj, Ax, state["sym"])
    sym, num = state["sym"], state["num"]
    L = Ax.shape[0]
    bounds = np.linspace(0, L, min(threads, L) + 1).astype(int)

    def work(lo, hi):
        if lo == hi:
            return
        if w.name == "cfg4":
            O.refactor(w.Ai, w.Aj, Ax[lo:hi], sym, num[lo:hi])
            O.solve_with_numeric(sym, num[lo:hi], b[lo:hi])
            O.solve_with_numeric(sym, num[lo:hi], b[lo:hi], transpose=True)
        else:
            O.solve_with_numeric(sym, num[lo:hi], b[lo:hi])
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(int(bounds[i]), int(bounds[i + 1]))) for i in range(len(bounds) - 1)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0


def cpu_sample_size(w: Workload) -> int:
    return {"cfg5": 8192, "cfg2": 16384, "cfg4": 16, "cfg3": 1}[w.name]


def config_of(w: Workload, world: int, flushed: bool) -> dict:
    """The same dict in both arms (the driver compares them)."""
    return {"workload": f"{w.name}: {w.desc}", "n_lhs_per_gpu": w.L, "n_col": w.n, "n_nz": w.nz, "n_rhs": w.r,
            "parallelism": f"n_lhs sharded over {world} GPU(s), no data-path collective",
            "l2": "L2 flushed (256 MB write) before every timed step" if flushed
                  else "inputs larger than L2 (no flush needed)"}


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU implementation of the path (restated KLU: the
    real one cannot be built here) on all host cores.  cfg5 / cfg2: the whole batch of the config
    per step; cfg4 / cfg3: a bounded sample, stated in cpu_baseline.sample."""
    if rank != 0:
        return
    w = Workload(args.workload, args.lhs)
    cores = len(os.sched_getaffinity(0))
    S = w.L if w.name in ("cfg5", "cfg2") else min(w.L, cpu_sample_size(w))
    r_keep = w.r
    if w.name == "cfg3":
        w.r = 8  # bounded: 8 of the 256 right-hand sides per step, scaled below
    Ax, b = w.values(S), w.rhs(S)
    state: dict = {}
    for _ in range(args.warmup):
        cpu_step(w, Ax, b, cores, state)
    total = 0.0
    for _ in range(args.steps):
        total += cpu_step(w, Ax, b, cores, state)
    scale = r_keep / w.r
    w.r = r_keep
    value = S * args.steps / (total * scale)
    sample = f"{S} systems per step" + (f", 8 of {r_keep} RHS (time scaled x{scale:g})" if scale != 1 else "")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total * scale / args.steps * (w.L / S),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "c128" if w.dtype is np.complex128 else "f64", "data": "synthetic",
           "config": config_of(w, args.gpus, w.name == "cfg2"),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": sample + "; oracle/ restated KLU driven like klujax.cpp, "
                                               "batch ranges spread over host threads"},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def measure(K, name: str, lhs, steps: int, warmup: int, rank: int, world: int, dev, dist, n_check: int,
            e2e_steps: int = 0, want_cpu: bool = False) -> dict:
    """One workload on this rank's GPU: device-timed steps (CUDA events on the launching stream,
    max over ranks), the parity gate on a sample of the timed batch, optionally the end-to-end leg
    and the one-thread CPU baseline."""
    import torch
    from klujax_b200 import _cabi
    lib = _cabi.lib()
    w = Workload(name, lhs, rank)
    cx = w.dtype is np.complex128
    tdt = torch.complex128 if cx else torch.float64
    sfx = "c128" if cx else "f64"
    L, n, r, nz = w.L, w.n, w.r, w.nz

    # synthetic inputs, resident in HBM before anything is timed
    Ax_np, b_np = w.values(L, seed=rank), w.rhs(L, seed=rank)
    Ax_h = torch.from_numpy(Ax_np).pin_memory()
    b_h = torch.from_numpy(b_np).pin_memory()
    Ai_d = torch.from_numpy(w.Ai).to(dev)
    Aj_d = torch.from_numpy(w.Aj).to(dev)
    Ax_d, b_d = Ax_h.to(dev), b_h.to(dev)
    x_d = torch.empty_like(b_d)
    xt_d = None
    st_d = torch.zeros(L, dtype=torch.int32, device=dev)
    gr_d = torch.zeros(L, dtype=torch.float64, device=dev)
    sym = K.analyze(w.Ai, w.Aj, n)
    symh = C.c_uint64(int(sym.handle))
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    numeric = None
    if w.name in ("cfg5", "cfg2"):
        f = getattr(lib, f"klujax_cuda_solve_with_symbol_{sfx}")

        def step():
            _cabi.check(f(symh, L, n, r, nz, P(Ai_d), P(Aj_d), P(Ax_d), P(b_d), P(x_d), P(st_d), P(gr_d), sp))
        launches_per_step = 2  # coo_map_kernel + wide_kernel / small_kernel (plus one 4-byte memset)
        kernel = "wide_kernel<zc>" if w.name == "cfg5" else "small_kernel<zc>"
    elif w.name == "cfg4":
        numeric = K.factor(Ai_d, Aj_d, Ax_d, sym)
        nh = np.ascontiguousarray(numeric.handle)
        nout = np.zeros_like(nh)
        xt_d = torch.empty_like(b_d)
        f_re = getattr(lib, f"klujax_cuda_refactor_{sfx}")
        f_so = getattr(lib, f"klujax_cuda_solve_with_numeric_{sfx}")
        f_ts = getattr(lib, f"klujax_cuda_tsolve_with_numeric_{sfx}")

        def step():
            _cabi.check(f_re(symh, L, nz, P(Ai_d), P(Aj_d), P(Ax_d), _cabi.u64p(nh), _cabi.u64p(nout),
                             P(st_d), P(gr_d), sp))
            _cabi.check(f_so(symh, L, _cabi.u64p(nh), L, n, r, P(b_d), P(x_d), sp))
            _cabi.check(f_ts(symh, L, _cabi.u64p(nh), L, n, r, P(b_d), P(xt_d), sp))
        launches_per_step = 6 + 3 + 3
        kernel = "lv_prepare_sysmajor + lv_factor (refactor), panel_solve (solve, transpose solve)"
    else:  # cfg3
        numeric = K.factor(Ai_d, Aj_d, Ax_d, sym)
        nh = np.ascontiguousarray(numeric.handle)
        f_so = getattr(lib, f"klujax_cuda_solve_with_numeric_{sfx}")

        def step():
            _cabi.check(f_so(symh, L, _cabi.u64p(nh), L, n, r, P(b_d), P(x_d), sp))
        launches_per_step = 3
        kernel = "lv_solve_grouped (row-blocked multi-right-hand-side solve)"

    flush = None
    if w.name == "cfg2":  # working set fits in the 126 MB L2: evict it between timed steps
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step()  # seeds the pivot sequence (one-time host stage), not a timed step
    torch.cuda.synchronize()
    info = _cabi.symbolic_info(int(sym.handle), cx)
    for _ in range(warmup):
        step()
    barrier()
    clocks = ClockSampler(dev.index)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for s0, s1 in ev:
        if flush is not None:
            flush.fill_(1)
        s0.record(stream)
        step()
        s1.record(stream)
    barrier()
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ok = int(st_d.abs().max().item()) == 0
    res = {"name": w.name, "workload": w, "ms_per_step": ms_total / steps, "value": world * L * steps / (ms_total * 1e-3),
           "status_all_ok": ok, "clocks": clk, "schedule": info, "launches_per_step": launches_per_step,
           "kernel": kernel, "flushed": flush is not None}

    # ---- parity gate on a sample of the batch that was just timed (rank 0)
    if rank == 0 and n_check > 0:
        res["parity"] = parity_sample(w, Ax_np, b_np, x_d.cpu().numpy(), n_check,
                                      xt_d.cpu().numpy() if xt_d is not None else None)

    # ---- roofline of the dominant kernel (algorithmic bytes / its event-timed duration)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    nnz_lu = None
    if w.name in ("cfg3", "cfg4"):
        nnz_lu = info["nnz_l"] + info["nnz_u"] + n + info["nzoff"]
    per_sys = w.algorithmic_bytes_per_system(nnz_lu)
    achieved = per_sys * L / (ms_total / steps * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": f"{kernel} (whole step timed with CUDA events; the COO map and status kernels are "
                          "~2% of it per profiles/launches_cfg5_r02.csv)",
                "algorithmic_bytes_per_system": per_sys}
    if w.name == "cfg5":  # what actually binds (ncu, profiles/README.md): on-chip, not HBM
        roofline["binding_resource"] = ("shared-memory wavefronts ~68% of peak and issue slots ~51% busy "
                                        "(ncu --set full, profiles/): the on-chip pipes of the SM with six "
                                        "resident systems, not HBM")
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        rec = json.load(open(tr)).get(w.name)
        # DRAM bytes of one launch (ncu --set full) - only if the capture was taken with THIS library
        if rec and rec.get("lib_sha16") == lib_fingerprint():
            roofline["traffic"] = rec["bytes_per_system"] * L
            roofline["traffic_source"] = rec["source"]
        elif rec:
            roofline["traffic_note"] = (f"no ncu capture of this build ({rec['source']} measured "
                                        f"{rec['bytes_per_system']} B/system with build {rec.get('lib_sha16')})")
    res["roofline"] = roofline

    # ---- end to end through the public API, host buffers, copies inside the timed region
    if e2e_steps > 0:
        x_h = torch.empty((L, n, r), dtype=tdt).pin_memory()
        Ai_h, Aj_h = torch.from_numpy(w.Ai).pin_memory(), torch.from_numpy(w.Aj).pin_memory()

        def e2e_step():
            if w.name in ("cfg5", "cfg2"):
                # H2D + kernels + D2H of x (pipelined per chunk) + status check
                K.solve_with_symbol(Ai_h, Aj_h, Ax_h, b_h, sym, out=x_h)
                return
            elif w.name == "cfg4":
                num2 = K.refactor(Ai_h, Aj_h, Ax_h, numeric, sym)
                x = K.solve_with_numeric(num2, b_h, sym)
                K.tsolve_with_numeric(num2, b_h, sym)
            else:
                x = K.solve_with_numeric(numeric, b_h, sym)
            x_h.copy_(x, non_blocking=True)                               # D2H of the result
            torch.cuda.current_stream().synchronize()
        e2e_step()
        e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(e2e_steps):
            e2e_step()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item()) / e2e_steps
        h2d = (Ax_h.numel() + b_h.numel()) * w.es + 2 * nz * 4
        if w.name == "cfg3":
            h2d = b_h.numel() * w.es
        if w.name == "cfg4":
            h2d = Ax_h.numel() * w.es + 2 * b_h.numel() * w.es + 2 * nz * 4
        d2h = x_h.numel() * w.es + (4 * L if w.name != "cfg3" else 0)
        # the box's own ceiling for these bytes: plain pinned copies, both directions at once on two
        # streams, every rank at the same time (what the e2e step cannot beat)
        hb = torch.empty(int(h2d), dtype=torch.uint8).pin_memory()
        db = torch.empty(int(h2d), dtype=torch.uint8, device=dev)
        hb2 = torch.empty(int(d2h), dtype=torch.uint8).pin_memory()
        db2 = torch.empty(int(d2h), dtype=torch.uint8, device=dev)
        s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def copies():
            with torch.cuda.stream(s_up):
                db.copy_(hb, non_blocking=True)
            with torch.cuda.stream(s_dn):
                hb2.copy_(db2, non_blocking=True)
        copies()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        c0.record(stream)
        s_up.wait_stream(stream)
        s_dn.wait_stream(stream)
        for _ in range(reps):
            copies()
        stream.wait_stream(s_up)
        stream.wait_stream(s_dn)
        c1.record(stream)
        barrier()
        t = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        copy_ms = float(t.item()) / reps
        res["e2e"] = {"value": world * L / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": e2e_ms,
                      "pcie": {"copy_only_ms_per_step": copy_ms,
                               "h2d_gbs_per_gpu": h2d / (copy_ms * 1e-3) / 1e9,
                               "ceiling_systems_per_s": world * L / (copy_ms * 1e-3),
                               "frac_of_ceiling": copy_ms / e2e_ms,
                               "how": "pinned cudaMemcpyAsync of the same bytes, H2D and D2H on two streams, "
                                      "all ranks at once, max over ranks"}}
        del hb, db, hb2, db2

    # ---- one-thread CPU baseline on a bounded sample of the same workload
    if want_cpu and world == 1:
        S = min(L, cpu_sample_size(w))
        w2 = Workload(name, lhs)
        if w2.name == "cfg3":
            w2.r = 8
        Ax_s, b_s, st_cpu = w2.values(S), w2.rhs(S), {}
        dt, reps = 0.0, 0
        while dt < 10.0 and reps < 64:  # a bounded sample: ~10-15 s of single-thread work
            dt += cpu_step(w2, Ax_s, b_s, 1, st_cpu)
            reps += 1
        scale = w.r / w2.r
        res["cpu_baseline"] = {
            "value": S * reps / (dt * scale), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{reps} x {S} systems, 1 thread, {dt:.1f} s" + (f" for {w2.r} of {w.r} RHS (scaled)" if scale != 1 else "")
                      + "; oracle/ (restated KLU driven like klujax.cpp:431-451: per-system klu_factor+klu_solve)"}
    if numeric is not None:
        numeric.close()
    sym.close()
    del Ax_d, b_d, x_d, Ax_h, b_h
    torch.cuda.empty_cache()
    return res


def cfg1_latency(K, dev) -> dict:
    """cfg1: the reference's own CPU-runnable case (klujax.solve, f64, n=1000, one system): a
    latency figure, per call once the pattern is cached, checked against the oracle."""
    import torch
    from oracle import oracle as O
    Ai, Aj, Ax = synth.random_dd()
    b = np.random.default_rng(1).standard_normal((1000,))
    Ai_d, Aj_d = torch.from_numpy(Ai).to(dev), torch.from_numpy(Aj).to(dev)
    Ax_d, b_d = torch.from_numpy(Ax[0]).to(dev), torch.from_numpy(b).to(dev)
    for _ in range(4):  # first call: host analysis + seed factorization + schedule (pattern cache)
        x = K.solve(Ai_d, Aj_d, Ax_d, b_d)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        x = K.solve(Ai_d, Aj_d, Ax_d, b_d)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    xo = O.solve(Ai, Aj, Ax, b[None, :, None])[0, :, 0]
    t0 = time.perf_counter()
    O.solve(Ai, Aj, Ax, b[None, :, None])
    cpu_ms = (time.perf_counter() - t0) * 1e3
    return {"workload": "cfg1: klujax.solve float64, random diagonally dominant, n_col=1000, n_nz=4993, n_lhs=1 (latency)",
            "ms_per_call": ms, "cpu_oracle_ms_per_call": cpu_ms,
            "parity": {"max_relerr": float(np.abs(x.cpu().numpy() - xo).max() / np.abs(xo).max()), "n_checked": 1}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--lhs", type=int, default=None, help="systems per GPU (default: the config's)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other named configs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import klujax_b200 as K

    m = measure(K, args.workload, args.lhs, args.steps, args.warmup, rank, world, dev, dist, n_check=256,
                e2e_steps=max(args.e2e_steps, 10), want_cpu=not args.no_cpu_baseline)
    w = m["workload"]
    extra = None
    if rank == 0 and world == 1 and not args.no_extra and args.workload == "cfg5":
        # the other named configs, rank 0 only, fewer steps: driver-visible numbers + parity for each
        extra = {}
        for nm, st_, nc in (("cfg2", 10, 256), ("cfg4", 3, 4), ("cfg3", 2, 1)):
            try:
                e = measure(K, nm, None, st_, 3, 0, 1, dev, dist, n_check=nc)
                extra[nm] = {"workload": f"{nm}: {e['workload'].desc}", "ms_per_step": e["ms_per_step"],
                             "systems_per_s": e["value"], "steps": st_, "status_all_ok": e["status_all_ok"],
                             "roofline_frac": e["roofline"]["frac"], "achieved_gbs": e["roofline"]["achieved"],
                             "algorithmic_bytes_per_system": e["roofline"]["algorithmic_bytes_per_system"],
                             "kernel": e["kernel"], "parity": e.get("parity"), "l2_flushed": e["flushed"]}
            except Exception as ex:  # an extra workload never takes the headline down
                extra[nm] = {"error": f"{type(ex).__name__}: {ex}"}
        try:
            extra["cfg1"] = cfg1_latency(K, dev)
        except Exception as ex:
            extra["cfg1"] = {"error": f"{type(ex).__name__}: {ex}"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cx = w.dtype is np.complex128
    out = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "c128" if cx else "f64", "data": "synthetic",
           "config": config_of(w, world, m["flushed"]), "schedule": m["schedule"],
           "clocks": m["clocks"], "status_all_ok": m["status_all_ok"], "parity": m.get("parity"),
           "e2e": m.get("e2e"), "gpu_launches": m["launches_per_step"] * args.steps,
           "roofline": m["roofline"], "cpu_baseline": m.get("cpu_baseline"),
           "extra_workloads": extra, "build": {"lib_sha16": lib_fingerprint(), "numa": numa}}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
