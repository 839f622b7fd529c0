#!/usr/bin/env python
"""bench.py -- headline benchmark of the FE assembly hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    (CPU port of the reference path)

Workload (BASELINE.json config 3 / 5): neo-Hookean hex8, 3 dof/node, structured 192^3 elements PER GPU
(weak scaling: 1 GPU 192^3, 2 GPUs 384x192x192, 4 GPUs 384x384x192, 8 GPUs 384^3), Float64.  N > 1: the global grid
is partitioned with METIS (k-way on the dual graph of 8^3-element cells; --partition brick gives flat bricks), ghost
residual contributions travel over NVLink, the collective plane (NCCL) lives inside libfecb200.  One
"step" = what one Newton iteration asks of the path: assemble_vector!(residual) + residual(asm), then
assemble_stiffness!(stiffness) into the CSR values.  value = elements assembled per second over the
whole job (all ranks), inputs resident in HBM.  `e2e` is the same step driven through the C ABI with
pinned HOST buffers (H2D of Uu, D2H of the residual inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 for every rank; the library's host-side set-up (tiling, adjacency, DOF maps) is
# OpenMP code, so give each rank its share of the cores BEFORE libgomp is loaded (set-up only; nothing timed uses it)
if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "finiteelementcontainers.jl_b200"))

# algorithmic (compulsory) DRAM bytes per element, SURVEY.md 8(d) / DESIGN.md section 5
BYTES_RESIDUAL = 136.0     # conn 64 + X 24 + U 24 + R 24
BYTES_TANGENT = 2066.0     # conn 64 + X 24 + U 24 + CSR values 1954 (nnz/NE * 8)
BYTES_FUSED = 2090.0       # tangent + R 24 (conn/X/U shared with the residual)
BYTES_ACTION = 160.0
# FP64 flops per element counted from the kernels' instruction mix
FLOPS_RESIDUAL = 4.5e3      # Walsh form of the HEX8 residual (estimate from instruction counts: 7.0e3 of the quadrature loop - 1728 FMA + ~1000 add/mul)
FLOPS_TANGENT = 14.0e3      # fused k_mat2, Walsh form: thread-level (2 DFMA + DMUL + DADD) per element = 2*4500 + 1600 + 3430, ncu r02i
FLOPS_TANGENT_QLOOP = 34.0e3  # the same element matrix by the plain quadrature loop (k_mat2 before the Walsh form, ncu r01z)
BYTES_ZERO_FILL = 1954.0   # fill!(storage, 0) of the CSR values (Matrix.jl:39): NOT algorithmic (SURVEY 8d counts every
                           # output once); reported as `extra_bytes_per_element` when the kernel clears the idle buffer
NEO_PROPS = np.array([1e3, 10.0e6, 1.0e6])


def ncu_traffic(dbuf):
    """DRAM bytes per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum from the committed
    `ncu --set full` capture of this command (profiles/traffic.json names the capture).  None when no capture of the
    current kernel is on file -- never a stale constant."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    e = d.get("double_buffered" if dbuf else "single_buffer")
    return (e.get("dram_bytes_per_launch"), e.get("source")) if e else (None, None)


def ncu_onchip(dbuf):
    """On-chip pipe utilisation of the dominant kernel from the same committed capture (the kernel is bound by the
    L1/TEX LSU data pipe, not by DRAM or the FP64 pipe): {'lsu_data_pipe_pct', 'fp64_pipe_pct', ...} or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    e = json.load(open(p)).get("double_buffered" if dbuf else "single_buffer")
    return e.get("onchip") if e else None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled through NVML
    every 5 ms (falls back to the nvidia-smi query line if pynvml is unavailable)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._halt = index, [], set(), threading.Event()
        self.max_mhz = None
        # NVML is initialised HERE, before the timed region: on a fresh box `import pynvml` + nvmlInit take longer than
        # the 10 timed steps, and a sampler that starts late reports no clocks at all
        self._nv = self._h = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self._h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM))
            nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
            self._nv = nv
        except Exception:
            self._nv = self._h = None

    def _run_nvml(self):
        nv, h = self._nv, self._h
        if nv is None:
            raise RuntimeError("NVML unavailable")
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._halt.is_set():
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = int(get(h))
            for n, b in bits.items():
                if r & b:
                    self.reasons.add(n)
            self._halt.wait(0.005)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.1)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_min_mhz": float(min(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def grid_for(world):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]


METIS_CELL = 8   # METIS runs on the dual graph of 8^3-element cells (384^3 elements -> 48^3 cells)


def make_partition(F, n, rank, world, how, cell=METIS_CELL):
    """rank-local mesh + Partition of the (gx n) x (gy n) x (gz n) grid.  metis: METIS_PartGraphKway on the coarse-cell
    dual graph (computed by rank 0, broadcast), every rank materialises only its own ragged piece."""
    from fecb200.partition import metis_cell_partition, structured_brick_partition, structured_cell_partition
    g = grid_for(world)
    if how == "brick":   # the same builder with one n^3 cell per rank: rank = px + gx (py + gy pz)
        ix, iy, iz = np.meshgrid(np.arange(g[0]), np.arange(g[1]), np.arange(g[2]), indexing="ij")
        lm, part = structured_cell_partition(F, tuple(gi * n for gi in g), ix + g[0] * (iy + g[1] * iz), n, rank, h=1.0 / n)
        part.cell = n
        return lm, part
    import torch
    import torch.distributed as dist
    cell = cell if n % cell == 0 else next(c for c in (8, 6, 4, 3, 2, 1) if n % c == 0)
    cells = tuple(gi * n // cell for gi in g)
    cp = torch.zeros(cells, dtype=torch.int32, device="cuda")
    if rank == 0:
        cp.copy_(torch.from_numpy(metis_cell_partition(cells, world).astype(np.int32)))
    if dist.is_initialized() and world > 1:
        dist.broadcast(cp, src=0)
    lm, part = structured_cell_partition(F, tuple(gi * n for gi in g), cp.cpu().numpy(), cell, rank, h=1.0 / n)
    part.cell = cell
    return lm, part


def raw_state(X, n, rng=None, H=1.0):
    """raw-throughput state: u = 0.02 (sin 2 pi y, sin 2 pi z, sin 2 pi x) [+ U(-1e-3,1e-3) h], times
    4 y (1 - y) / H^2 so that it MEETS the Dirichlet data on the bottom / top faces: without the taper the fixed faces
    (u = 0) sit one element away from |u| = 0.02, the boundary layer of elements is sheared to det F <= 0 and its
    1/J terms dominate (and ill-condition) every norm -- round 1's state did that."""
    taper = 4.0 * (X[1] / H) * (1.0 - X[1] / H)
    U = 0.02 * np.stack([np.sin(2 * np.pi * X[1]), np.sin(2 * np.pi * X[2]), np.sin(2 * np.pi * X[0])])
    if rng is not None:
        U += rng.uniform(-1e-3, 1e-3, U.shape) / n
    return U * taper


def build_problem(F, n, rank, world, matrix_free=False, partition="metis", noise=True, mesh_part=None, height=None):
    """Rank-local neo-Hookean problem.  N = 1: StructuredMesh('hex', (0,0,0), (1,1,1), (n+1,)*3) with the
    BCs of BASELINE config 3.  N > 1: this rank's piece of the global grid (see fecb200.partition)."""
    verbose = bool(os.environ.get("FECB200_VERBOSE"))
    tick = [time.time()]
    laps = {}

    def lap(name):
        laps[name] = round(time.time() - tick[0], 3)
        if verbose and rank == 0:
            print(f"[bench setup] {name:28s} {time.time() - tick[0]:.3f} s", file=sys.stderr, flush=True)
        tick[0] = time.time()
    if mesh_part is not None:
        mesh, part = mesh_part
    elif world == 1:
        mesh = F.StructuredMesh("hex", (0., 0., 0.), (1., 1., 1.), (n + 1, n + 1, n + 1))
        part = None
    else:
        mesh, part = make_partition(F, n, rank, world, partition)
    lap("mesh")
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange, q_type="GaussLegendre", q_degree=2)
    u = F.VectorFunction(V, "displ")
    lap("function space")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=False, matrix_free=matrix_free,
                                  device=int(os.environ.get("LOCAL_RANK", 0)))
    lap("SparseMatrixAssembler")
    zero = lambda X, t: np.zeros(X.shape[0])
    pull = lambda X, t: np.full(X.shape[0], 0.1 * t)
    dbcs = [F.DirichletBC(c, zero, nodeset_name="bottom") for c in u.names()]
    dbcs += [F.DirichletBC("displ_x", zero, nodeset_name="top"), F.DirichletBC("displ_z", zero, nodeset_name="top"),
             F.DirichletBC("displ_y", pull, nodeset_name="top")]
    p = F.create_parameters(mesh, asm, F.NeoHookean(F.ThreeDimensional()), NEO_PROPS, dirichlet_bcs=dbcs,
                            times=F.TimeStepper(0.0, 1.0, 10))
    lap("create_parameters")
    if part is not None:
        part.attach(asm)
        lap("partition attach")
    X = np.asarray(mesh.nodal_coords)
    H = float(grid_for(world)[1]) if height is None else float(height)      # y-extent of the GLOBAL grid (element size 1/n)
    U = raw_state(X, n, np.random.default_rng(42 + rank) if noise else None, H)
    Uu = np.ascontiguousarray(U.reshape(-1, order="F")[asm.dof.unknown_dofs - 1])
    lap("initial state")
    asm._setup_laps = laps
    return mesh, asm, p, Uu, part


def bind_to_gpu_numa_node(index):
    """CPU affinity of this process = the cores NVML reports as local to the GPU (one process per GPU)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index), (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def prefer_gpu_numa_memory(index):
    """Best effort: make this process allocate (and therefore pin) host memory on the NUMA node the GPU hangs off
    (set_mempolicy(MPOL_PREFERRED, node)); the pinned staging buffers of the e2e path are allocated afterwards.
    Returns the node or None (no NUMA information, e.g. a virtualised host)."""
    try:
        import ctypes
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:      # 00000000:17:00.0 -> 0000:17:00.0
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        mask = (ctypes.c_ulong * 2)(0, 0)
        mask[node // 64] = 1 << (node % 64)
        libc = ctypes.CDLL(None, use_errno=True)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238
        if libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), 129) != 0:
            return None
        return node
    except Exception:
        return None


def _rel(a, b):
    import torch
    d = float(torch.linalg.vector_norm(a - b))
    s = float(torch.linalg.vector_norm(b))
    return d / s if s > 0 else d


def check_single(F, asm, p, dUu, stream):
    """Self-check of the VERY problem that is timed (N = 1, full size): the assembled tangent applied on the device
    against the matrix-free action, the fused kernel's residual against the stand-alone residual kernel, and a
    checksum of the values.  (Parity against the oracle at size: tests/test_gpu_at_size.py.)"""
    import torch
    N = len(dUu)
    with torch.cuda.stream(stream):
        v = torch.rand(N, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
        one = torch.ones(N, dtype=torch.float64, device="cuda")
        Rf, Rv, Kv, Av, K1 = (torch.empty(N, dtype=torch.float64, device="cuda") for _ in range(5))
    stream.synchronize()
    F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, dUu, p)
    F.residual(asm, Rf)
    F.matrix_multiply(asm, v, Kv)
    F.matrix_multiply(asm, one, K1)
    F.assemble_matrix_free_action(asm, F.stiffness_action, dUu, v, p)
    F.hvp(asm, v, Av)
    F.assemble_vector(asm, F.residual, dUu, p)
    F.residual(asm, Rv)
    stream.synchronize()
    torch.cuda.synchronize()
    e_kv, e_r = _rel(Kv, Av), _rel(Rf, Rv)
    finite = bool(torch.isfinite(Kv).all() and torch.isfinite(Rf).all())
    return {"Kv_vs_matrix_free_action": e_kv, "fused_R_vs_residual_kernel": e_r, "sum_nzval": float(K1.sum()),
            "norm_R": float(torch.linalg.vector_norm(Rf)), "tol": 1e-11, "ok": bool(finite and e_kv < 1e-11 and e_r < 1e-11)}


def check_partitioned(F, rank, world, partition, n_c=48):
    """N > 1: the partitioned path (same partitioner, library communicator, fused peer-memory halo) against a SERIAL
    handle on the same global mesh, small enough to fit one GPU (n_c^3 elements per rank; 96^3 global at N = 8), under
    this very launch: owned residual rows, K v and the row sums K 1 on owned rows."""
    import torch
    import torch.distributed as dist
    from fecb200 import _lib
    from fecb200._lib import check, lib
    g = grid_for(world)
    # ---- partitioned
    mesh, asm, p, Uu, part = build_problem(F, n_c, rank, world, partition=partition, noise=False)
    part.comm_init(asm)
    part.enable_peer_scatter(asm)
    h = asm._require()
    dUu = torch.from_numpy(Uu).cuda()
    N = len(Uu)
    tmp = torch.empty(N, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    F.residual(asm, tmp)                       # flush: every rank's residual field is zero ...
    part.barrier_on_stream()                   # ... before anyone scatters
    asm.set_matrix_double_buffer(True)
    for _ in range(2):                         # second pass lands in kernel-cleared storage
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, dUu, p)
        part.halo_sum_residual(asm)
        R_field = F.full_field(asm, "residual").reshape(-1, 3)[:part.n_owned_nodes].copy()
        F.residual(asm, tmp)
        part.barrier_on_stream()
    # ---- serial, same global mesh, on this rank's GPU
    E = tuple(gi * n_c for gi in g)
    gmesh = F.StructuredMesh("hex", (0., 0., 0.), tuple(e / n_c for e in E), tuple(e + 1 for e in E))
    _, gasm, gp, gUu, _ = build_problem(F, n_c, 0, 1, noise=False, mesh_part=(gmesh, None), height=g[1])
    F.assemble_vector_and_stiffness(gasm, F.residual, F.stiffness, gUu, gp)
    Rg = F.full_field(gasm, "residual").reshape(-1, 3)
    l2g = part.local_to_global - 1
    own = l2g[:part.n_owned_nodes]
    e_R = float(np.abs(R_field - Rg[own]).max() / np.abs(Rg).max())
    # K v on owned rows: local unknown -> global unknown through the dof ids
    nown = C_int64_value(lib.fecb200_owned_length, h)
    ud_l = asm.dof.unknown_dofs[:nown] - 1
    gdof = 3 * l2g[ud_l // 3] + ud_l % 3
    ug = gasm.dof.dof_to_unknown[gdof] - 1
    assert (ug >= 0).all()
    rng = np.random.default_rng(7)
    vg = rng.uniform(0, 1, gasm.sizes()[2])
    # local v: owned AND ghost entries from the global vector (the library refreshes ghosts anyway)
    ud_all = asm.dof.unknown_dofs - 1
    ug_all = gasm.dof.dof_to_unknown[3 * l2g[ud_all // 3] + ud_all % 3] - 1
    errs = {}
    for name, xg in (("Kv", vg), ("rowsum", np.ones_like(vg))):
        yl = F.matrix_multiply(asm, np.ascontiguousarray(xg[ug_all]))[:nown]
        yg = F.matrix_multiply(gasm, xg)
        errs[name] = float(np.abs(yl - yg[ug]).max() / np.abs(yg).max())
    gasm.close()
    t = torch.tensor([e_R, errs["Kv"], errs["rowsum"]], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    torch.cuda.synchronize()
    asm.close()
    e_R, e_kv, e_rs = (float(x) for x in t.tolist())
    return {"against": f"serial assembly of the same {E[0]}x{E[1]}x{E[2]} mesh on every rank's own GPU", "partition": partition,
            "owned_R_vs_serial": e_R, "Kv_owned_rows_vs_serial": e_kv, "rowsum_owned_rows_vs_serial": e_rs, "tol": 1e-11,
            "ok": bool(max(e_R, e_kv, e_rs) < 1e-11)}


def C_int64_value(fn, h):
    import ctypes as C
    from fecb200._lib import check
    v = C.c_int64()
    check(fn(h, C.byref(v)))
    return v.value


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    numa_node = None
    if world > 1:
        bind_to_gpu_numa_node(local)   # pinned staging buffers of the e2e path land next to this rank's GPU
        numa_node = prefer_gpu_numa_memory(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import fecb200 as F
    from fecb200 import _lib
    from fecb200._lib import check, lib

    n = args.n
    t0 = time.time()
    mesh, asm, p, Uu_h, part = build_problem(F, n, rank, world, partition=args.partition)
    setup_s = time.time() - t0
    ne_local = mesh.element_conns["block_1"].shape[1] if part is None else part.n_owned_elements
    h = asm._require()
    stream = torch.cuda.Stream()
    check(lib.fecb200_set_stream(h, stream.cuda_stream))
    N = len(Uu_h)
    with torch.cuda.stream(stream):
        dUu = torch.from_numpy(Uu_h).cuda()
        dR = torch.empty_like(dUu)
    hUu = torch.from_numpy(Uu_h).pin_memory()
    hR = torch.empty(N, dtype=torch.float64).pin_memory()
    stream.synchronize()

    dbuf = not args.single_buffer
    if dbuf:
        # two CSR value arrays: each assembly fills the one the previous assembly's kernel cleared (TMA bulk stores
        # under the element kernel) instead of running a stand-alone fill!(storage, 0) pass every step
        asm.set_matrix_double_buffer(True)
    peer = part is not None and not args.nccl_halo
    if part is not None:
        # the library's own NCCL communicator (fecb200_comm_init): torch.distributed only carries the 128-byte id
        part.comm_init(asm)
        if peer:
            # fused halo: ghost-node REDs go straight into the owner's residual over NVLink peer memory
            part.enable_peer_scatter(asm)
        stream.synchronize()

    def halo():
        if part is not None:
            part.halo_sum_residual(asm, stream)   # peer mode: a stream-ordered barrier; else pack / ncclSend+Recv / add

    def pre():
        if peer:
            part.barrier_on_stream(stream)        # every rank has read and re-zeroed its residual

    def step_device():
        # assemble_vector!(residual) + assemble_stiffness!(stiffness) at the same Uu, as solve! does
        # (src/Solvers.jl:133-140), through the fused entry point; then the ghost->owner sum and residual(asm)
        pre()
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, dUu, p)
        halo()
        F.residual(asm, dR)

    def step_unfused():
        pre()
        F.assemble_vector(asm, F.residual, dUu, p)
        halo()
        F.residual(asm, dR)
        F.assemble_stiffness(asm, F.stiffness, dUu, p)

    def step_e2e():
        pre()
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, hUu, p)  # H2D of Uu inside
        halo()
        F.residual(asm, hR)                             # D2H of the residual inside (synchronous)

    def barrier():
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if peer:
        F.residual(asm, dR)            # flush: every rank's residual field is zero before the first scatter
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = asm.launch_count()
    ms = timed(step_device, args.steps)
    launches = asm.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ne_total = ne_local * world if part is None else part.n_global_elements
    value = ne_total * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers; copies run on the library's copy streams
    # (fecb200_set_async) and are all complete at the closing fecb200_synchronize inside the timed region
    check(lib.fecb200_set_async(h, 1))
    for _ in range(2):
        step_e2e()
    check(lib.fecb200_synchronize(h))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    check(lib.fecb200_synchronize(h))      # host blocks until the last D2H landed in pinned memory ...
    e1.record(stream)                      # ... so this event closes the region including the copy streams
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    check(lib.fecb200_set_async(h, 0))
    e2e_value = ne_total * args.steps / (ms_e2e * 1e-3)

    # ---- partition statistics (N > 1): imbalance, neighbours, ghost fraction -- gathered over the ranks
    pstats = None
    if part is not None:
        mine = torch.tensor([part.n_owned_elements, getattr(part, "n_halo_elements", 0), part.n_owned_nodes,
                             len(part.local_to_global) - part.n_owned_nodes, len(part.neighbors),
                             sum(len(v) for v in part.send.values())], device="cuda", dtype=torch.float64)
        allp = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        A = torch.stack(allp).cpu().numpy()
        pstats = {"owned_elements_max_over_mean": round(float(A[:, 0].max() / A[:, 0].mean()), 4),
                  "owned_elements_per_rank": [int(x) for x in A[:, 0]],
                  "halo_elements_per_rank": [int(x) for x in A[:, 1]],
                  "neighbours_per_rank": [int(x) for x in A[:, 4]],
                  "ghost_node_fraction_max": round(float((A[:, 3] / (A[:, 2] + A[:, 3])).max()), 5),
                  "residual_halo_nodes_per_rank": [int(x) for x in A[:, 5]]}

    # ---- correctness of what was just timed
    if args.no_check:
        chk = None
    elif world == 1:
        chk = check_single(F, asm, p, dUu, stream)
    else:
        chk = check_partitioned(F, rank, world, args.partition)

    out = None
    if rank == 0:
        hbm, peak_src, peaks = load_peaks()
        # ---- per-operation device timings (CUDA events on the launching stream) + roofline of the dominant kernel
        def op_ms(fn, reps=5):
            for _ in range(2):
                fn()
            return timed(fn, reps) / reps if world == 1 else None

        roof, ops, cfgs = {}, {}, None
        if world == 1:
            Vu = torch.rand(N, dtype=torch.float64, device="cuda")
            t_unf = op_ms(step_unfused)
            t_sb = None
            if dbuf:
                asm.set_matrix_double_buffer(False)
                t_sb = op_ms(step_device)
                asm.set_matrix_double_buffer(True)
            t_res = op_ms(lambda: F.assemble_vector(asm, F.residual, dUu, p))
            t_tan = op_ms(lambda: F.assemble_stiffness(asm, F.stiffness, dUu, p))
            t_act = op_ms(lambda: F.assemble_matrix_free_action(asm, F.stiffness_action, dUu, Vu, p))
            # the operator application of the device CG on the assembled CSR values (SpMV), next to the matrix-free one
            Yu = torch.empty_like(Vu)
            t_spmv = op_ms(lambda: F.matrix_multiply(asm, Vu, Yu))
            # dominant kernel alone: events recorded by the library around the element kernel launch
            check(lib.fecb200_enable_timing(h, 1))
            import ctypes as C

            def kernel_ms(fn):
                xs = []
                for _ in range(6):
                    fn()
                    f = C.c_float()
                    check(lib.fecb200_last_kernel_ms(h, C.byref(f)))
                    xs.append(f.value)
                return float(np.mean(xs[1:]))
            k_tan = kernel_ms(lambda: F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, dUu, p))
            k_res = kernel_ms(lambda: F.assemble_vector(asm, F.residual, dUu, p))
            k_act = kernel_ms(lambda: F.assemble_matrix_free_action(asm, F.stiffness_action, dUu, Vu, p))
            check(lib.fecb200_enable_timing(h, 0))
            # FP64 roof measured here with cuBLAS DGEMM (MEASURED_PEAKS.json has no FP64 entry)
            a = torch.randn(6144, 6144, dtype=torch.float64, device="cuda")
            torch.mm(a, a)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.mm(a, a); torch.mm(a, a); e1.record(); torch.cuda.synchronize()
            fp64_peak = 2 * 2 * 6144 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
            del a
            # compulsory bytes, every output counted once.  fused = conn 64 + X 24 + U 24 + CSR values 1954 + R 24
            ach = BYTES_FUSED * ne_local / (k_tan * 1e-3) / 1e9
            tf = FLOPS_TANGENT * ne_local / (k_tan * 1e-3) / 1e12
            traffic, traffic_src = ncu_traffic(dbuf)
            onchip = ncu_onchip(dbuf)
            roof = {"bound": "hbm",
                    "binding_roof": ("scatter: L2 RED-sector rate / L1-LSU data pipe (on chip)" if onchip and onchip.get("lsu_data_pipe_pct", 0) > 100 * max(tf / fp64_peak, ach / hbm)
                                     else "fp64" if tf / fp64_peak > ach / hbm else "hbm"),
                    "onchip_ncu": onchip,
                    "kernel": "fused residual + tangent -> CSR (" + os.environ.get("FECB200_MAT_KERNEL", "default") + ")",
                    "achieved": round(ach, 1), "peak": hbm, "peak_source": peak_src, "unit": "GB/s", "frac": round(ach / hbm, 4),
                    "algorithmic_bytes_per_element": BYTES_FUSED,
                    "extra_bytes_per_element": BYTES_ZERO_FILL if dbuf else 0.0,
                    "extra_bytes_note": "in-kernel clear of the idle CSR value array (fill!(storage, 0), Matrix.jl:39); not algorithmic" if dbuf else None,
                    "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": round(k_tan, 4),
                    "fp64_flops_per_element": FLOPS_TANGENT, "fp64_flops_per_element_quadrature_loop": FLOPS_TANGENT_QLOOP,
                    "fp64_achieved_tflops": round(tf, 2),
                    "fp64_peak_tflops_dgemm_measured": round(fp64_peak, 1), "fp64_frac": round(tf / fp64_peak, 4),
                    "residual_kernel_ms": round(k_res, 4),
                    "residual_frac_hbm": round(BYTES_RESIDUAL * ne_local / (k_res * 1e-3) / 1e9 / hbm, 4),
                    "residual_frac_fp64": round(FLOPS_RESIDUAL * ne_local / (k_res * 1e-3) / 1e12 / fp64_peak, 4),
                    "action_kernel_ms": round(k_act, 4),
                    "action_frac_hbm": round(BYTES_ACTION * ne_local / (k_act * 1e-3) / 1e9 / hbm, 4)}
            ops = {"residual_elements_per_s": round(ne_local / (t_res * 1e-3), 1),
                   "tangent_elements_per_s": round(ne_local / (t_tan * 1e-3), 1),
                   "action_elements_per_s": round(ne_local / (t_act * 1e-3), 1),
                   "unfused_step_ms": round(t_unf, 4), "unfused_step_elements_per_s": round(ne_local / (t_unf * 1e-3), 1),
                   "residual_ms": round(t_res, 4), "tangent_ms": round(t_tan, 4), "action_ms": round(t_act, 4),
                   "csr_spmv_ms": round(t_spmv, 4)}
            if t_sb is not None:
                ops["single_buffer_step_ms"] = round(t_sb, 4)   # same step with cudaMemset of the CSR values instead
        cpu = cpu_baseline(args.cpu_n) if (world == 1 and not args.no_cpu) else None
        par = "single GPU"
        if world > 1:
            par = (("METIS k-way on the dual graph of %d^3-element cells" % getattr(part, "cell", METIS_CELL)) if args.partition == "metis"
                   else "structured bricks") + f" x{world}, owned/ghost nodes, halo-element block for local Jacobian rows, residual halo = " + \
                  ("fused peer-memory REDs over NVLink + 2 library NCCL barriers" if peer else "library NCCL send/recv")
        out = {
            "metric": "assembled elements/s (residual + Jacobian), neo-Hookean hex8 FP64",
            "value": round(value, 1), "unit": "elements/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"neohookean_hex8_{n}^3_per_gpu residual+tangent(CSR) per Newton iteration",
                       "elements_per_gpu": int(ne_local), "elements_total": int(ne_total), "dofs_per_gpu": int(len(asm.dof)),
                       "csr_nnz_per_gpu": int(asm.pattern()[2].shape[0]) if args.report_nnz else None,
                       "parallelism": par, "partition": pstats,
                       "l2": "inputs and outputs larger than L2 (CSR values 13.8 GB at 192^3); no flush needed",
                       "csr_values": "double-buffered, idle buffer cleared inside the element kernel" if dbuf else "single buffer + memset per step",
                       "setup_s": round(setup_s, 1),
                       "setup_breakdown_s": getattr(asm, "_setup_laps", None),
                       "setup_note": "mesh = host mesh generation (numpy), create_parameters = the library set-up (device-side tile / adjacency / CSR-offset build, host DOF maps) + first-use CUDA module load, initial state = the benchmark's synthetic displacement field"},
            "e2e": {"value": round(e2e_value, 1), "unit": "elements/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                    "h2d_bytes_per_step": int(N * 8), "d2h_bytes_per_step": int(N * 8),
                    "pcie_GBs_per_rank": round(2 * N * 8 / (ms_e2e / args.steps * 1e-3) / 1e9, 1),
                    "pcie_GBs_all_ranks": round(world * 2 * N * 8 / (ms_e2e / args.steps * 1e-3) / 1e9, 1),
                    "host_numa_node_rank0": numa_node, "host_cpus_visible": len(os.sched_getaffinity(0))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "check": chk,
        }
        if roof:
            out["roofline"] = roof
        if ops:
            out["ops"] = ops
        if cpu:
            out["cpu_baseline"] = cpu
    if world > 1:
        dist.barrier()
    asm.close()
    if rank == 0 and world == 1 and not args.no_configs:
        # BASELINE configs 2 and 4 (kernel timings with their own roofline fractions), same process, after the main run
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_configs
            out["configs"] = [bench_configs.poisson(128), bench_configs.j2(64)]
        except Exception as e:  # never lose the headline line to a side measurement
            out["configs"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def _cpu_problem(n):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fec_oracle as O
    import fec_oracle_clib as OC
    m = O.structured_mesh("hex", (0., 0., 0.), (1., 1., 1.), (n + 1,) * 3)
    X = m["coords"]
    U = 0.02 * np.stack([np.sin(2 * np.pi * X[1]), np.sin(2 * np.pi * X[2]), np.sin(2 * np.pi * X[0])])
    cp = OC.CProblem(m["conn"], X, O.ref_fe_tables("HEX8", "gauss2"), "neo", 3, NEO_PROPS)
    return OC, cp, U.reshape(-1, order="F"), m


def cpu_step_factory(n):
    """One reference-style CPU step: assemble_vector! + assemble_stiffness! (COO) + stiffness(asm) (sparse!)."""
    OC, cp, Uf, m = _cpu_problem(n)
    nthreads = OC.max_threads()
    Is, Js = cp.pattern()
    ndof = 3 * m["coords"].shape[1]
    ws = OC.SparseWorkspace(len(Is), ndof)
    slots = np.arange(1, len(Is) + 1, dtype=np.int64)
    coo = np.empty(len(Is))

    def step():
        cp.assemble_vector(Uf, nthreads=nthreads)
        cp.assemble_matrix_coo(Uf, 2, nthreads=nthreads, out=coo)
        ws.sparse_csc(Is, Js, slots, coo)

    return step, cp.ne, nthreads


def cpu_baseline(n):
    step, ne, nthreads = cpu_step_factory(n)
    step()
    t0 = time.perf_counter()
    reps = 0
    while reps < 2 or (time.perf_counter() - t0 < 8.0 and reps < 20):
        step()
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {"value": round(ne / dt, 1), "unit": "elements/s", "cores": nthreads, "kind": "port",
            "sample": f"neo-Hookean hex8 {n}^3 ({ne} elements, NOT the 192^3 of the GPU arm: the reference's COO bookkeeping "
                      f"does not fit a host there, BASELINE.md section 4): residual + COO tangent + sparse!, {reps} reps, reported per "
                      "element; C/OpenMP port of the reference CPU path with analytic tangents (faster than the Julia AD path; "
                      "Julia is not installable here)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Julia and
    cannot be installed in this image (no julia, no network), so this times the C/OpenMP port of the
    same algorithm (oracle/fec_oracle_c.c) with all host threads, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n = args.cpu_n
    step, ne, nthreads = cpu_step_factory(n)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = ne * args.steps / dt
    sample = (f"neo-Hookean hex8 {n}^3 ({ne} elements) per step, not the GPU arm's 192^3 per GPU (the reference's COO arrays do not "
              "fit a host there; per-element rate reported): residual + COO tangent + sparse!, C/OpenMP port, all host threads")
    print(json.dumps({
        "impl": "reference", "metric": "assembled elements/s (residual + Jacobian), neo-Hookean hex8 FP64",
        "value": round(v, 1), "unit": "elements/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"neohookean_hex8_{args.n}^3_per_gpu residual+tangent(CSR) per Newton iteration",
                   "cpu_sample": sample},
        "cpu_baseline": {"value": round(v, 1), "unit": "elements/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 1), "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=int(os.environ.get("FECB200_BENCH_N", 192)), help="elements per axis per GPU")
    ap.add_argument("--cpu-n", type=int, default=64, help="elements per axis of the bounded CPU sample (BASELINE.md section 4: 64^3)")
    ap.add_argument("--partition", default="metis", choices=["metis", "brick"], help="N > 1: METIS k-way on the coarse-cell dual graph (BASELINE config 5) or flat bricks")
    ap.add_argument("--no-check", action="store_true", help="skip the correctness block")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE config 2 / 4 kernel timings")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--nccl-halo", action="store_true", help="N > 1: pack / NCCL send-recv / unpack instead of the fused peer-memory scatter")
    ap.add_argument("--report-nnz", action="store_true")
    ap.add_argument("--single-buffer", action="store_true", help="one CSR value array + cudaMemset per step instead of the double-buffered in-kernel clear")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
