#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 assembly back end (contract: see the task statement / DESIGN.md section 6).

Workload (config.workload): BASELINE.json configs[1] -- LinearElastic p=2 tetrahedra, 55^3 hexes split in 6 -> 998 250 tet10
elements, 1 367 631 nodes, implicit stiffness + residual assembled to CSR (requires_geometry_update forced to 1 so that T is
computed).  One "step" = one full assembly (K values in CSR order + T) of the whole mesh.
  value : elements/s with the state already in HBM (CUDA events around K steps, max over ranks)
  e2e   : the same assembly through the reference-facing plug-in function with HOST numpy state in and HOST K values / T out
  roofline : SURVEY.md 8(d)'s algorithmic bytes per element x elements / the device-timed step / the measured HBM peak -- the WHOLE
             step (all kernels of one assembly), with the per-kernel figures under `kernels`; `traffic` is the ncu dram__bytes of
             the same kernels from the capture named in `traffic_source` (null when no capture of this round exists)
  roofline_fp64 : executed fp64 work of the element kernel against the DFMA peak measured in this run (builder-measured)
  cpu_baseline : the CPU restatement of the reference algorithm (oracle, -O3 -ffast-math) on a bounded sample, 1 core
  explicit / explicit5a / explicit5b : the metric's second half -- explicit central-difference DOF-updates/s for config 3
             (NeoHookean hex27) and config 5 (MooneyRivlin hex8 / hex27), 200^3 elements per GPU (weak); explicit3_strong (N > 1):
             config 3's 200^3 mesh split over the ranks
  hiorder : config 4 (EM_108 hex64 Newton-step assembly) against the measured DMMA peak;  poisson : config 1 (hex125 Laplacian)
N > 1 (torchrun): weak scaling, every rank assembles its own slab (+1 halo layer of elements so the CSR rows of the nodes it
owns are complete; no data-path collective) and emits the CSR row block it owns with global column numbers (values: a slice of
V; columns and the all_gather'ed offsets belong to the pattern and are built once); the explicit lines exchange interface forces
over NCCL every step (after the element forces; the order that overlaps the exchange with the interior elements is timed beside it).
`--impl reference` times the reference algorithm (oracle port) on all host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=55, help="hexes per edge before the 6-tet split (55 -> 998250 tet10)")
    ap.add_argument("--explicit-n", type=int, default=200, help="hexes per edge of the hex27 explicit line (per GPU)")
    ap.add_argument("--explicit-steps", type=int, default=20)
    ap.add_argument("--no-explicit", action="store_true")
    ap.add_argument("--no-hiorder", action="store_true")
    ap.add_argument("--hiorder-n", type=int, default=24, help="hexes per edge of the p=3 electro-mechanical Newton-step line (config 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--no-poisson", action="store_true")
    return ap.parse_args()


MU, NU = 1.0e5, 0.3
LAMB = 2.0 * MU * NU / (1.0 - 2.0 * NU)
METRIC = "elements assembled/s (K+residual, fp64)"


class Clocks(object):
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
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
            if len(f) < 7:
                continue
            try:
                if float(f[2]) < 250.0:   # idle sample (power draw in W), not part of the load
                    mx.append(float(f[1]))
                    continue
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class _Obj(object):
    pass


def reference_objects(points, elements, Bases, Jm, AllGauss, ndim, nvar, recompute=False):
    """Duck-typed stand-ins for (fem_solver, function_space, formulation, mesh, material) as the reference's wrappers read them."""
    fs, fo, me, so = _Obj(), _Obj(), _Obj(), _Obj()
    fs.Bases, fs.Jm, fs.AllGauss = Bases, Jm, AllGauss
    fo.ndim, fo.nvar, fo.fields = ndim, nvar, "mechanics"
    me.points, me.elements, me.nelem = points, elements, elements.shape[0]
    me.ChangeType = lambda: None
    so.recompute_sparsity_pattern, so.squeeze_sparsity_pattern, so.requires_geometry_update = recompute, False, True
    return so, fs, fo, me


# =================================================================================================== reference arm
def run_reference(args, rank, world):
    """The reference algorithm (CPU restatement; the Fastor/CBLAS build of Florence cannot be produced) on all host
    cores, in the SAME mode as the B200 arm (recompute_sparsity_pattern=False: CSR values through the precomputed slot maps,
    _MassIntegrand_.h:142-151 / SparseAssemblyNative.h:32-45): contiguous element blocks over a thread pool (the native call
    releases the GIL), every block returns its own (V, T) as _LowLevelAssembly_Par_ does (_LowLevelAssembly_.py:87-105) and the
    parent sums them (Assembly.py:1000-1041).  The sparsity pattern and slot maps are built once, outside the timed region, like
    fl_pattern_build in the B200 arm.  (The COO + scipy COO->CSR variant of the same pool is 2.9x slower on 8 cores.)"""
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from florence_b200 import mesh as flmesh
    from oracle import oracle as orc
    orc.build()
    cores = len(os.sched_getaffinity(0))
    n = args.n
    # bounded sample: a slab of the same mesh, ~cores * 15k elements per step (every block holds a private nnz-long V)
    per_core = 15000
    nz = max(1, min(n, int(round(min(cores * per_core, 300000) / (6.0 * n * n)))))
    pts, els = flmesh.box_tet_mesh(n, n, nz, p=2, lengths=(1.0, 1.0, float(nz) / n))
    pts, els = pts.numpy(), els.numpy().astype(np.uint64)
    Bases, Jm, AG = flmesh.tables("tet", 2)
    rng = np.random.default_rng(0)
    x = pts + 1e-3 * rng.uniform(-1, 1, pts.shape)
    prm = orc.params(mu=MU, lamb=LAMB)
    nelem, nnode = els.shape[0], pts.shape[0]
    blocks = np.array_split(np.arange(nelem), cores)
    pat = orc.sparsity_pattern(els, nnode, 3)          # indices, indptr, data_local_indices, data_global_indices (one-off)
    nnz = pat[0].shape[0]

    def work(b):
        out = (np.zeros(nnz), np.zeros(nnode * 3))
        return orc.assemble_implicit(pts, els, x, None, Jm, AG, 3, 6, 1, prm, 10, mode="csr", pattern=pat,
                                     elem_range=(int(b[0]), int(b[-1]) + 1), out=out, fast=True)

    def step():
        with ThreadPoolExecutor(cores) as ex:
            res = list(ex.map(work, blocks))
        V, T = res[0]
        for r in res[1:]:
            V += r[0]
            T += r[1]
        return V, T

    for _ in range(min(args.warmup, 1)):
        step()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    val = nelem * steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "elements/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "tet10 LinearElastic implicit K(CSR)+T, %d^3 hexes x6 (reference arm: bounded slab sample)" % n},
            "cpu_baseline": {"value": val, "unit": "elements/s", "cores": cores, "kind": "port",
                             "sample": "%d tet10 elements per step (slab %dx%dx%d of the %d^3 mesh), %d thread-pool blocks, CSR slot-map mode, per-block (V, T) "
                                       "summed on the parent; pattern + slot maps built once outside the timed region" % (nelem, n, n, nz, n, cores)},
            "e2e": {"value": val, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "steps_note": "CPU arm: at most 3 timed steps and 1 warm-up of a bounded slab sample whatever --steps/--warmup ask for "
                          "(one step is ~1.4 s of all host cores; elements/s is size-independent here)",
            "gpu_launches": 0}
    print(json.dumps(line))


# =================================================================================================== B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    from florence_b200 import assembly, backend, mesh as flmesh, partition, time_integrator
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    launches = 0
    hbm_peak, peak_src = measured_peaks()
    # ------------------------------------------------------------------ implicit tet10 (config 2)
    n = args.n
    halo = 1 if (world > 1 and rank < world - 1) else 0
    pts, els = flmesh.box_tet_mesh(n, n, n + halo, p=2, lengths=(1.0, 1.0, float(n + halo) / n), device=dev)
    pts[:, 2] += float(rank)  # slab `rank` of a box stacked along z
    Bases, Jm, AG = flmesh.tables("tet", 2)
    nelem_local, nnode = els.shape[0], pts.shape[0]
    nelem_owned = 6 * n * n * n
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    x = pts + 1e-3 * (2.0 * torch.rand(pts.shape, dtype=torch.float64, device=dev, generator=gen) - 1.0)
    h = backend.AssemblyHandle(pts, els, Jm, AG, Bases, device=dev)
    mat = backend.make_material(10, 0.0, mu=MU, lamb=LAMB)
    nnz = h.build_pattern(3)
    V = torch.empty(nnz, dtype=torch.float64, device=dev)
    T = torch.empty(nnode * 3, dtype=torch.float64, device=dev)
    ndof = 30
    # N > 1: the rows this rank owns (node planes of its own slab; the bottom plane belongs to the rank below, whose halo layer
    # completes it), their global numbering and the position of the block in the global CSR (all_gather + exclusive scan, once)
    row_info = None
    if world > 1:
        plane = (2 * n + 1) ** 2
        p0, p1 = (1 if rank > 0 else 0), 2 * n + 1
        owned = torch.arange(p0 * plane, p1 * plane, dtype=torch.int32, device=dev)
        node_map = torch.arange(nnode, dtype=torch.int64, device=dev) + rank * 2 * n * plane
        ipb, colsb, valsb = h.row_block(3, V, owned, node_map)
        ro, no, tr, tn = partition.global_row_offsets(ipb.numel() - 1, int(ipb[-1]), device=dev)
        row_info = {"row_block_emitted": True, "rows": int(ipb.numel() - 1), "nnz_block": int(ipb[-1]), "row_offset": ro, "nnz_offset": no,
                    "global_rows": tr, "global_nnz": tn,
                    "values": "zero-copy slice of V (owned node planes are a contiguous row range)" if valsb.data_ptr() != 0 and
                              valsb.untyped_storage().data_ptr() == V.untyped_storage().data_ptr() else "compacted copy"}

    def step():
        h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V, T))
        if world > 1:
            h.row_block(3, V, owned, node_map)

    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(1.0)   # nvidia-smi needs about a second before its first sample
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1))
    barrier()
    launches += 3 * args.steps
    value = nelem_owned * world * args.steps / (ms * 1e-3)
    # per-kernel timing (events inside the library, same stream)
    h.set_timing(True)
    kms = []
    for _ in range(args.steps):
        h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V, T))
        kms.append(h.get_timing())
    h.set_timing(False)
    launches += 3 * args.steps
    # hold the same load for ~0.5 s so the 20 ms nvidia-smi sampler sees the clocks the timed region ran at
    t_hold = time.perf_counter()
    nhold = 0
    while time.perf_counter() - t_hold < 0.5:
        h.assemble_implicit(x, None, mat, 0, True, mode="csr", out=(V, T))
        torch.cuda.synchronize()
        nhold += 1
    launches += 3 * nhold
    clk = clocks.stop() if rank == 0 else None
    kms = np.array(kms)
    k_elem, k_csr, k_T = kms.mean(0)

    # ------------------------------------------------------------------ e2e through the reference-facing plug-in (host in / host out)
    x_host = x.cpu().numpy()
    pts_host, els_host = pts.cpu().numpy(), els.cpu().numpy().astype(np.uint64)
    so, fs, fo, me = reference_objects(pts_host, els_host, Bases, Jm, AG, 3, 3, recompute=False)
    material = _Obj(); material.mu, material.lamb, material.rho, material.mtype = MU, LAMB, 1.0, "LinearElastic"
    # pre-seed the handle cache with the handle that already holds this mesh (same mesh arrays -> no re-upload)
    assembly._handle_cache[(assembly._array_key(pts_host), assembly._array_key(els_host), assembly._array_key(Jm))] = h
    func = assembly._LowLevelAssemblyDF__LinearElastic_
    e2e_steps = max(2, min(args.steps, 5))

    def time_e2e():
        for _ in range(2):
            out = func(so, fs, fo, me, material, x_host, None)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = func(so, fs, fo, me, material, x_host, None)
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0), out
    # headline e2e: the plug-in's DEFAULT mode.  Every call returns arrays the caller owns, like the reference; large results are
    # leased pinned buffers that go back to a free list when the caller's last reference is gone (assembly._lease), so a loop that
    # drops K before it re-assembles -- this one, and a Newton loop -- pays no 2.8 GB allocation and no host-side copy.
    e2e_s, (Vh, Th) = time_e2e()
    assert np.array_equal(Vh, V.cpu().numpy()), "host path and device path disagree"
    h2d = x_host.nbytes
    d2h = Vh.nbytes + Th.nbytes
    # worst case of the same mode: the caller keeps EVERY result alive, so after three leases the results are ordinary pageable
    # arrays (staged copy-out, pages populated while the device works)
    held = [func(so, fs, fo, me, material, x_host, None) for _ in range(3)]
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        held.append(func(so, fs, fo, me, material, x_host, None))
    torch.cuda.synchronize()
    fresh_s = max_over_ranks(time.perf_counter() - t0)
    assert np.array_equal(held[-1][0], Vh) and not np.shares_memory(held[-1][0], held[-2][0]) and not np.shares_memory(held[-1][0], Vh)
    del held
    launches += 3 * (2 * e2e_steps + 5)     # (e2e_steps + 2) + (3 + e2e_steps) plug-in calls, three kernels each
    e2e_value = nelem_owned * world * e2e_steps / e2e_s

    # ------------------------------------------------------------------ the same Newton-iteration assembly with K kept on the device
    # (SURVEY.md 8f.1): host state in -> plug-in with device_out -> Dirichlet reduction on the device -> only the reduced residual F_b
    # returns to the host; K_b stays in HBM for a device solver.  Bottom face clamped, top face displaced.
    from florence_b200 import boundary
    zc = pts_host[:, 2]
    flags = np.full((pts_host.shape[0], 3), np.nan)
    flags[np.isclose(zc, zc.min())] = 0.0
    flags[np.isclose(zc, zc.max()), 2] = 0.01
    bc = boundary.DeviceBoundaryCondition.from_flags(h, flags)
    def newton_assembly():
        Vd, Td = func(so, fs, fo, me, material, x_host, None, device_out=True)
        Kb, Fb, _ = bc.ApplyDirichletGetReducedMatrices(Vd, Td, bc.applied_dirichlet, LoadFactor=1.0)
        return Kb, assembly._to_host(Fb, "Fb")
    for _ in range(2):
        Kb, Fb_host = newton_assembly()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        Kb, Fb_host = newton_assembly()
    torch.cuda.synchronize()
    res_s = max_over_ranks(time.perf_counter() - t0)
    launches += 4 * (e2e_steps + 2)
    resident = {"value": nelem_owned * world * e2e_steps / res_s, "unit": "elements/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(Fb_host.nbytes), "nnz_reduced": int(bc.nnz_b), "n_free_dofs": int(bc.n_in),
                "api": "plug-in (device_out=True) + boundary.DeviceBoundaryCondition.ApplyDirichletGetReducedMatrices; K_b stays on the device"}
    del Kb

    # ------------------------------------------------------------------ roofline
    nnode_per_elem = nnode / float(nelem_local)
    # B = 8 npe + (nnode/nelem)(2*8*d) + (nnode/nelem) 8 nvar + 4 ndof^2 + 8 nnz/nelem
    B_alg = 8 * 10 + nnode_per_elem * (2 * 8 * 3) + nnode_per_elem * 8 * 3 + 4 * ndof * ndof + 8.0 * nnz / nelem_local
    Fl_ref = 8 * (10 * 9 * 10 + 100 + 150 + (2 * 30 * 6 + 2 * 30)) + 8 * (2 * 36 * 30 + 2 * 6 * 900 + 2 * 900 + 25 * 100)
    dfma = backend.measure_fp64_peak(False, 20000)
    launches += 4
    traffic, traffic_src = {}, None
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            traffic = json.load(open(tr))
            traffic_src = traffic.get("_source")
        except Exception:
            traffic = {}
    if not traffic_src:
        traffic = {}          # only a capture that names its commit / command is quoted
    k_step = k_elem + k_csr + k_T
    step_ms = ms / args.steps
    # what each kernel has to move in THIS design (not SURVEY's whole-path figure): the element kernel reads the state and writes
    # the K_e scratch (along the Morton curve) and the per-element tractions; the reduction reads the scratch and its packed plan
    # (per node visit: flat index + 4-bit slot fields, 24 B for <= 32 neighbours, 56 B for <= 96) and writes the CSR values
    B_elem = 8 * 10 / 2 + nnode_per_elem * 48 + 8.0 * ndof * ndof + 8.0 * ndof
    B_csr = 8.0 * ndof * ndof + 36.0 * 10 + 8.0 * nnz / nelem_local
    kernels = [
        {"kernel": "implicit_iso_warp_kernel<10,8>", "kernel_ms": float(k_elem), "design_bytes_per_element": B_elem,
         "design_GBps": B_elem * nelem_local / (k_elem * 1e-3) / 1e9, "traffic": traffic.get("implicit_iso_warp_kernel_bytes_per_launch"),
         "share_of_step": float(k_elem / k_step)},
        {"kernel": "csr_gather_curve_kernel<10>", "kernel_ms": float(k_csr), "design_bytes_per_element": B_csr,
         "design_GBps": B_csr * nelem_local / (k_csr * 1e-3) / 1e9, "traffic": traffic.get("csr_gather_curve_kernel_bytes_per_launch"),
         "share_of_step": float(k_csr / k_step)},
        {"kernel": "gather_traction_kernel", "kernel_ms": float(k_T), "traffic": traffic.get("gather_traction_kernel_bytes_per_launch"),
         "share_of_step": float(k_T / k_step),
         "note": "runs on a side stream beside the CSR reduction (0.10 ms alone): kernel_ms is what is left of it after the reduction has finished"}]
    for kk in kernels:
        if "design_GBps" in kk:
            kk["design_frac_of_hbm_peak"] = kk["design_GBps"] / hbm_peak
    tsum = [kk["traffic"] for kk in kernels]
    ach = B_alg * nelem_local / (step_ms * 1e-3) / 1e9
    dom = max(kernels, key=lambda kk: kk["kernel_ms"])
    roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
            "traffic": float(sum(tsum)) if all(t is not None for t in tsum) else None, "traffic_source": traffic_src,
            "scope": "whole step: SURVEY.md 8(d) algorithmic bytes per element x elements / device-timed ms_per_step (all kernels of one assembly)",
            "algorithmic_bytes_per_element": B_alg, "algorithmic_bytes_per_step": B_alg * nelem_local, "step_ms": step_ms,
            "kernel": dom["kernel"], "kernel_ms": dom["kernel_ms"],
            "dominant_kernel_frac": B_alg * nelem_local / (dom["kernel_ms"] * 1e-3) / 1e9 / hbm_peak,
            "dominant_kernel_note": "the same algorithmic bytes charged to the dominant kernel alone",
            "peak_source": peak_src, "kernels": kernels}
    # executed fp64 work of the element kernel (LinearElastic takes the isotropic constant-tangent path):
    # kinematics 8 gp x (18 npe + ~120) + spatial gradients 8*10*9 + S_ab 8 gp x 100 node pairs x 9 (the warp-autonomous kernel
    # computes every block, no symmetry) + traction, FMA = 2; + the combination lamb S + mu S^T + mu tr I (~30 per node pair)
    Fl_exec = 2.0 * (8 * (18 * 10 + 120) + 8 * 10 * 9 + 8 * 100 * 9 + 8 * 10 * 9) + 100 * 30
    roof64 = {"bound": "fp64", "achieved": Fl_exec * nelem_local / (k_elem * 1e-3) / 1e12, "peak": dfma, "unit": "TFLOP/s",
              "frac": Fl_exec * nelem_local / (k_elem * 1e-3) / 1e12 / dfma, "flops_per_element_executed": Fl_exec,
              "kernel": "implicit_iso_warp_kernel<10,8>", "kernel_ms": float(k_elem),
              "reference_count": {"flops_per_element": Fl_ref, "equivalent_tflops": Fl_ref * nelem_local / (k_elem * 1e-3) / 1e12,
                                  "note": "what the reference's dense dgemm formulation would need for the same result (SURVEY.md 8d)"},
              "peak_source": "builder-measured in this run (fl_measure_fp64_peak, register-resident DFMA loop); MEASURED_PEAKS.json has no fp64 entry"}
    cfg = {"workload": "tet10 LinearElastic implicit K(CSR)+T, %d^3 hexes x6 = %d elements per GPU" % (n, nelem_owned),
           "nelem_per_gpu": nelem_owned, "halo_elements": nelem_local - nelem_owned, "nnode_per_gpu": nnode, "nnz_per_gpu": nnz,
           "mode": "CSR (recompute_sparsity_pattern=False), requires_geometry_update=1", "parallelism": "element slabs x%d" % world,
           "l2": "inputs+scratch (%.1f GB per step) larger than L2, no flush needed" % ((nelem_local * 900 * 8 * 2 + nnz * 8) / 1e9)}
    if row_info:
        cfg.update(row_info)
    line = {"metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "elements/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "florence_b200.assembly._LowLevelAssemblyDF__LinearElastic_ (host numpy in, host numpy out)",
                    "host_results": "default mode: every call returns arrays the caller owns (leased pinned buffers, returned to a free "
                                    "list by a weakref finaliser when the last reference incl. views is gone; never written while referenced)",
                    "all_results_held_value": nelem_owned * world * e2e_steps / fresh_s,
                    "all_results_held_note": "same mode when the caller keeps every result alive: beyond three leases the results are "
                                             "pageable arrays (chunked D2H, multi-threaded copy-out, pages populated during the device work)"},
            "e2e_device_resident": resident,
            "roofline": roof, "roofline_fp64": roof64}

    # ------------------------------------------------------------------ cpu baseline (rank 0, N=1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build()
        ns = 200000
        sub = els_host[:ns]
        t0 = time.perf_counter()
        orc.assemble_implicit(pts_host, sub, x_host, None, Jm, AG, 3, 6, 1, orc.params(mu=MU, lamb=LAMB), 10, mode="coo", fast=True)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns / dt, "unit": "elements/s", "cores": 1, "kind": "port",
                                "sample": "first %d elements of the same mesh, COO triplets + T, oracle -O3 -ffast-math, 1 thread (%.1f s)" % (ns, dt)}
    del V, T, Vh, Th
    h.close()
    assembly._handle_cache.clear()
    assembly._pinned.clear()
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ explicit central-difference lines (configs 3 and 5)
    def explicit_line(p, matnum, matname, prm, rho, nxy, nz_local, nsteps, scaling, label):
        """One explicit line: order-p hexahedra, nxy x nxy x nz_local elements on this rank, slabs stacked along z."""
        part = partition.slab_partition_hex(nxy, nxy, nz_local, p, rank, world, device=dev,
                                            lengths_per_rank=(1.0, 1.0, float(nz_local) / nxy))
        if world > 1:
            part.interface_first()      # interface elements first, so that the overlapped order can be timed on the same partition
        Bs, Jh, AGh = flmesh.tables("hex", p)
        hh = backend.AssemblyHandle(part.points, part.elements, Jh, AGh, Bs, device=dev)
        mat_n = backend.make_material(matnum, rho, **prm)
        ex = partition.InterfaceExchange(part, 3, dev, handle=hh) if world > 1 else None
        hx = 1.0 / nxy
        lam_eff = prm["lamb"] + 2 * (prm.get("mu", 0.0) + prm.get("mu1", 0.0) + prm.get("mu2", 0.0))
        dt_x = 0.2 * (hx / p) / np.sqrt(lam_eff / rho)
        nn = part.points.shape[0]
        fixed = torch.zeros(nn * 3, dtype=torch.uint8, device=dev)
        # a smooth perturbation that is a function of the position: interface nodes start identically on both owners
        x0 = part.points + 0.02 * (hx / p) * torch.sin(1000.0 * part.points)

        def make(exch, overlap=False):
            integ = time_integrator.ExplicitStructuralDynamicIntegrator(hh, mat_n, rho=rho, exchange=exch, overlap=overlap)
            integ.initialise(part.points, None, fixed, dt_x)
            integ.Eulerx.copy_(x0.reshape(-1)); integ.internal_force(integ.Eulerx.view(nn, 3), out=integ.T)
            integ.step(3, 2)
            return integ

        def timed(integ):
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            st = integ.step(nsteps, 5)
            a1.record()
            torch.cuda.synchronize()
            return max_over_ranks(a0.elapsed_time(a1)), st
        integ = make(ex)           # the integrator's default order: all element forces, then the exchange
        ems, st = timed(integ)
        out = {}
        if world > 1:
            # the same local work without any exchange (interface nodes then miss the neighbours' forces: timing only), and with the
            # exchange forced into either order: what the interface exchange costs per step, and what the overlap hides
            ems_ov, _ = timed(make(ex, overlap=True))
            ems_none, _ = timed(make(None))
            out["exchange"] = {"mode": "not overlapped (default)",
                               "ms_per_step_overlapped": ems_ov / nsteps, "ms_per_step_not_overlapped": ems / nsteps,
                               "ms_per_step_no_exchange": ems_none / nsteps, "exchange_cost_ms_per_step": (ems - ems_none) / nsteps,
                               "interface_bytes_per_step": ex.bytes_per_exchange(), "interface_elements": int(part.n_interface_elements or 0)}
        NX = p * nxy + 1
        nz_total = nz_local * world
        ndof_global = 3 * NX * NX * (p * nz_total + 1)
        hh.set_timing(True)
        hh.assemble_explicit(integ.Eulerx.view(nn, 3), None, mat_n, 0)
        t_el, _, t_g = hh.get_timing()
        hh.set_timing(False)
        nel = nxy * nxy * nz_local
        npe = (p + 1) ** 3
        ng = npe
        B_x = 8 * npe + (nn / nel) * (2 * 8 * 3) + (nn / nel) * 8 * 3
        matfl = 350.0 if matnum == 1 else 450.0
        Fl_x = ng * (10 * 9 * npe + 100 + 500 + (2 * 3 * npe * 6 + 2 * 3 * npe))
        # useful work of the regrouped algebra: Jacobian GEMM (3 ng x npe x 6) + traction GEMM (npe x 3 ng x 3), FMA = 2, + the
        # kinematics / material law per Gauss point; "executed" adds the padding of the 8x8x4 DMMA tiles (NE elements per batch)
        Fl_useful = 2.0 * (3 * ng) * npe * 6 + 2.0 * npe * (3 * ng) * 3 + ng * matfl
        NE = 8 if p == 2 else 16
        t1 = ((3 * ng + 7) // 8) * ((npe + 3) // 4) * (6 * NE // 8)
        t3 = ((npe + 7) // 8) * ((3 * ng + 3) // 4) * (3 * NE // 8) * (2 if p == 1 else 1)   # hex8: half the warps of the traction GEMM run an empty m-tile
        Fl_exec = (t1 + t3) / float(NE) * 512.0 + ng * matfl
        kname = "explicit_elements_mma_kernel<%s,%d,%d,%d>" % (matname, npe, ng, NE)
        out.update({"metric": "explicit DOF-updates/s", "value": ndof_global * nsteps / (ems * 1e-3), "unit": "DOF-updates/s",
                    "elements_per_s": nel * world * nsteps / (ems * 1e-3), "ms_per_step": ems / nsteps, "steps": nsteps,
                    "scaling": scaling, "status": int(st), "blew_up": bool(st),
                    "config": {"workload": label, "ndof": ndof_global, "elements_per_gpu": nel,
                               "interface_bytes_per_step": 0 if ex is None else ex.bytes_per_exchange()},
                    "roofline": {"bound": "hbm", "achieved": B_x * nel / (t_el * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": B_x * nel / (t_el * 1e-3) / 1e9 / hbm_peak, "kernel": kname, "kernel_ms": t_el, "gather_ms": t_g,
                                 "note": "fp64-pipe bound (SURVEY.md 8d): see roofline_fp64"},
                    "roofline_fp64": {"achieved": Fl_useful * nel / (t_el * 1e-3) / 1e12, "peak": dfma, "unit": "TFLOP/s",
                                      "frac": Fl_useful * nel / (t_el * 1e-3) / 1e12 / dfma, "flops_per_element_useful": Fl_useful,
                                      "flops_per_element_executed_incl_tile_padding": Fl_exec,
                                      "frac_executed": Fl_exec * nel / (t_el * 1e-3) / 1e12 / dfma,
                                      "peak_source": "builder-measured in this run (fl_measure_fp64_peak)",
                                      "reference_count": {"flops_per_element": Fl_x, "equivalent_tflops": Fl_x * nel / (t_el * 1e-3) / 1e12}}})
        hh.close()
        torch.cuda.empty_cache()
        return out, (2 * (nsteps + 3) + 4) * (3 if world > 1 else 1)

    if not args.no_explicit:
        ne = args.explicit_n
        line["explicit"], k = explicit_line(2, 1, "NeoHookean", dict(mu=4.0e5, lamb=2.0e6), 1100.0, ne, ne, args.explicit_steps, "weak",
                                            "config 3 shape: hex27 NeoHookean explicit central difference, %d^3 elements per GPU" % ne)
        launches += k
        if world > 1 and ne % world == 0:
            line["explicit3_strong"], k = explicit_line(2, 1, "NeoHookean", dict(mu=4.0e5, lamb=2.0e6), 1100.0, ne, ne // world,
                                                        args.explicit_steps, "strong",
                                                        "config 3: hex27 NeoHookean, %d^3 elements in total, %d z-layers per GPU" % (ne, ne // world))
            launches += k
    if not args.no_explicit and not args.no_config5:
        ne = args.explicit_n
        mu5, nu5 = 1.0e6, 0.495        # car_crash_analysis constants
        prm5 = dict(mu1=mu5, mu2=0.0, lamb=2.0 * mu5 * nu5 / (1.0 - 2.0 * nu5))
        line["explicit5a"], k = explicit_line(1, 2, "MooneyRivlin", prm5, 8000.0, ne, ne, args.explicit_steps, "weak",
                                              "config 5a: hex8 MooneyRivlin explicit, %d^3 elements per GPU stacked in z" % ne)
        launches += k
        line["explicit5b"], k = explicit_line(2, 2, "MooneyRivlin", prm5, 8000.0, ne, ne, args.explicit_steps, "weak",
                                              "config 5b: hex27 MooneyRivlin explicit, %d^3 elements per GPU stacked in z" % ne)
        launches += k
    # ------------------------------------------------------------------ config 4: EM_108 p=3 hex Newton-step assembly (DMMA local K)
    if not args.no_hiorder:
        n4 = args.hiorder_n
        p4, e4 = flmesh.box_hex_mesh(n4, n4, n4, p=3, device=dev)
        B4, J4, A4 = flmesh.tables("hex", 3)
        x4 = flmesh.perturbed_state(p4, 1.0 / (3 * n4), 0.02, seed=11)
        gen4 = torch.Generator(device=dev); gen4.manual_seed(5)
        phi4 = 9.0e3 * p4[:, 2] + 10.0 * (2.0 * torch.rand(p4.shape[0], dtype=torch.float64, device=dev, generator=gen4) - 1.0)
        h4 = backend.AssemblyHandle(p4, e4, J4, A4, B4, device=dev)
        nnz4 = h4.build_pattern(4)
        mu4 = 5.0e4
        mat4 = backend.make_material(8, 1200.0, mu1=mu4, mu2=mu4, lamb=2.0 * mu4 * 0.4 / (1.0 - 0.8), eps_2=4.0 * 8.8541e-12)
        V4 = torch.empty(nnz4, dtype=torch.float64, device=dev)
        T4 = torch.empty(p4.shape[0] * 4, dtype=torch.float64, device=dev)
        for _ in range(2):
            h4.assemble_implicit(x4, phi4, mat4, 1, True, mode="csr", out=(V4, T4))
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = 3
        b0.record()
        for _ in range(nrep):
            h4.assemble_implicit(x4, phi4, mat4, 1, True, mode="csr", out=(V4, T4))
        b1.record()
        torch.cuda.synchronize()
        hms = max_over_ranks(b0.elapsed_time(b1)) / nrep
        h4.set_timing(True)
        h4.assemble_implicit(x4, phi4, mat4, 1, True, mode="csr", out=(V4, T4))
        t4 = h4.get_timing()
        h4.set_timing(False)
        launches += 3 * (nrep + 3)
        dmma = backend.measure_fp64_peak(True, 20000)
        launches += 4
        nel4 = e4.shape[0]
        # executed tensor-core work per element: 6 off-diagonal dof pairs (i < j) x 64 tiles + 4 diagonal pairs x 36 upper-triangular
        # tiles (the mirror tiles of a symmetric K^{ii} are skipped), x 48 k-steps, DMMAs of 256 FMAs
        fl_exec = (6 * 64 + 4 * 36) * 48 * 512.0
        fl_ref = 64 * (2 * 81 * 256 + 2 * 9 * 65536 + 2 * 65536 + 25 * 4096)   # SURVEY.md 8(d) implicit term, reference-algorithm count
        line["hiorder"] = {"metric": "elements assembled/s (K+residual, fp64)", "value": nel4 * world / (hms * 1e-3), "unit": "elements/s",
                           "ms_per_step": hms,
                           "config": {"workload": "hex64 (p=3) IsotropicElectroMechanics_108 Newton-step K(CSR)+T, %d^3 elements per GPU" % n4,
                                      "ndof_per_element": 256, "nnz_per_gpu": nnz4},
                           "roofline": {"bound": "tensor", "achieved": fl_exec * nel4 / (t4[0] * 1e-3) / 1e12, "peak": dmma, "unit": "TFLOP/s",
                                        "frac": fl_exec * nel4 / (t4[0] * 1e-3) / 1e12 / dmma,
                                        "frac_whole_step": fl_exec * nel4 / (hms * 1e-3) / 1e12 / dmma,
                                        "kernel": "implicit_mma_prologue_kernel<EM_108,64,64> + implicit_mma_gemm_kernel<4,64,64,12,0>",
                                        "kernel_ms": t4[0], "csr_reduction_kernel": "csr_gather_hex64_kernel<4>", "csr_reduction_ms": t4[1],
                                        "reference_count": {"flops_per_element": fl_ref,
                                                            "equivalent_tflops_whole_step": fl_ref * nel4 / (hms * 1e-3) / 1e12,
                                                            "note": "what the reference's dense dgemm formulation would execute for the same K (SURVEY.md 8d)"},
                                        "peak_source": "builder-measured in this run (fl_measure_fp64_peak, mma.sync.m8n8k4.f64 loop); "
                                                       "MEASURED_PEAKS.json has no fp64 entry",
                                        "flops_per_element_executed_on_tensor_cores": fl_exec}}
        del V4, T4
        h4.close()
        torch.cuda.empty_cache()
    # ------------------------------------------------------------------ config 1: Poisson, p = 4 hexahedra (hex125)
    if not args.no_poisson:
        B1, J1, A1 = flmesh.tables("hex", 4)
        pois = {}
        for n1 in (6, 16):
            p1, e1_ = flmesh.box_hex_mesh(n1, n1, n1, p=4, device=dev)
            h1 = backend.AssemblyHandle(p1, e1_, J1, A1, B1, device=dev)
            nnz1 = h1.build_pattern(1)
            etens = -2.35 * np.eye(3)
            for _ in range(3):
                h1.assemble_laplacian(etens, True, mode="csr")
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nrep = 5
            c0.record()
            for _ in range(nrep):
                h1.assemble_laplacian(etens, True, mode="csr")
            c1.record()
            torch.cuda.synchronize()
            pms = max_over_ranks(c0.elapsed_time(c1)) / nrep
            h1.set_timing(True)
            h1.assemble_laplacian(etens, True, mode="csr")
            tk = h1.get_timing()
            h1.set_timing(False)
            launches += 2 * (nrep + 4)
            nel1 = e1_.shape[0]
            # upper triangle, npe^2/2 * ng * 6 flops = 5.9 Mflop per hex125 element
            fl1 = 125 * 125 / 2.0 * 125 * 6
            pois["%d^3" % n1] = {"elements": nel1, "nnz": nnz1, "ms_per_assembly": pms, "elements_per_s": nel1 * world / (pms * 1e-3),
                                 "element_kernel_ms": tk[0], "csr_reduction_ms": tk[1],
                                 "fp64_TFLOPs_reference_count": fl1 * nel1 / (tk[0] * 1e-3) / 1e12,
                                 "fp64_frac_reference_count": fl1 * nel1 / (tk[0] * 1e-3) / 1e12 / dfma}
            h1.close()
        line["poisson"] = {"metric": "elements assembled/s (K, fp64)", "config": {"workload": "config 1: Poisson, hex125 (p=4), K(CSR); "
                           "6^3 is the reference's simple_laplace mesh, 16^3 the same element at a size that fills the GPU"},
                           "value": pois["6^3"]["elements_per_s"], "unit": "elements/s", "sizes": pois,
                           "peak_source": "builder-measured DFMA peak of this run"}
        torch.cuda.empty_cache()
    line["gpu_launches"] = int(launches)
    if rank == 0:
        print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
