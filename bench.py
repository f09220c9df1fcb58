#!/usr/bin/env python
"""bench.py -- mesh-extraction benchmark (BASELINE.json metric: mesh-extraction s and
vertices/s vs host CPU; % HBM roofline).

A "step" is one full polyhedral-complex mesh extraction (skeleton -> 33 hyperplane
subdivisions -> faces) of one trilinear SDF network, planar (-f) path.  At N ranks every
rank extracts its own copy of the object (independent objects, no data-path collective):
weak scaling; `value` = mesh vertices extracted by all ranks / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload small_sphere|small_torus|large_random]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "tropical-nerf.pytorch_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "mesh_extraction_vertices_per_s"
UNIT = "vertices/s"


# ---------------------------------------------------------------------------------------
# workloads (synthetic: fitted-analytic-SDF fixtures or seeded random weights)
# ---------------------------------------------------------------------------------------
def load_workload(name):
    """Plain arrays of one network: dict(levels, n_feat, log2_T, n_min, per_level_scale,
    num_layers, num_hidden, table, mlp, marks, eps, scale, describe)."""
    if name in ("small_sphere", "small_torus"):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
        nl = int(g["net_num_layers"])
        mlp = np.concatenate([np.concatenate([g[f"net_w{i}"].reshape(-1), g[f"net_b{i}"].reshape(-1)])
                              for i in range(nl)]).astype(np.float32)
        shape = "sphere" if "sphere" in name else "torus"
        return dict(levels=int(g["net_levels"]), n_feat=int(g["net_n_feat"]), log2_T=int(g["net_log2_T"]),
                    n_min=int(g["net_n_min"]), per_level_scale=float(g["net_per_level_scale"]),
                    num_layers=nl, num_hidden=int(g["net_num_hidden"]), table=g["net_table"], mlp=mlp,
                    marks=g["net_marks"], eps=float(g["net_eps"]), scale=float(g["net_scale"]),
                    describe=f"small HashGrid+MLP (L=4 F=2 T=19 r=2..32, MLP 8-16-16-2) briefly fitted to an "
                             f"analytic {shape} SDF, planar (-f) extraction, marks grid {len(g['net_marks'])}^3")
    if name.split("_")[0] in ("medium", "large") and name.split("_")[1] in ("sphere", "torus"):
        return fitted_workload(*name.split("_")[:2])
    if name.endswith("_random"):
        from tropical.stanford.model import Net
        import torch
        size = name.split("_")[0]
        r_min, r_max = {"small": (2, 32), "medium": (4, 64), "large": (8, 128)}[size]
        torch.manual_seed(0)
        net = Net(num_layers=3, num_hidden=16, levels=4, r_min=r_min, r_max=r_max, T=19)
        with torch.no_grad():  # random-init weights with an amplitude that gives a surface
            net.enc.module.params.uniform_(-1.0, 1.0)
        mlp = np.concatenate([np.concatenate([fc.weight.detach().numpy().reshape(-1),
                                              fc.bias.detach().numpy().reshape(-1)]) for fc in net.fc])
        return dict(levels=4, n_feat=2, log2_T=19, n_min=r_min, per_level_scale=float(net.enc.b),
                    num_layers=3, num_hidden=16, table=net.enc.module.params.detach().numpy(),
                    mlp=mlp.astype(np.float32), marks=net.enc.marks.numpy(), eps=1e-4, scale=1.0,
                    describe=f"{size} HashGrid+MLP (r={r_min}..{r_max}), random-init weights (table U(-1,1)), "
                             f"planar (-f) extraction, marks grid {len(net.enc.marks)}^3")
    raise SystemExit(f"unknown workload {name}")


def fitted_workload(size, shape, steps=800, batch=1 << 16):
    """medium / large HashGrid+MLP briefly fitted to an analytic SDF (BASELINE.json: datasets
    and checkpoints are unavailable offline).  Workload GENERATION only: a plain torch
    autograd fit (pure-torch hash-grid gather), run once per box and cached so that both
    bench arms and every rank see identical weights."""
    import tempfile
    import torch
    from tropical.stanford.model import Net
    r_min, r_max = {"medium": (4, 64), "large": (8, 128)}[size]
    cache = os.path.join(tempfile.gettempdir(), f"tnb_workload_{size}_{shape}_{steps}.npz")
    torch.manual_seed(0)
    net = Net(num_layers=3, num_hidden=16, levels=4, r_min=r_min, r_max=r_max, T=19)
    describe = (f"{size} HashGrid+MLP (L=4 F=2 T=19 r={r_min}..{r_max}, MLP 8-16-16-2) briefly fitted "
                f"({steps} Adam steps) to an analytic {shape} SDF, planar (-f) extraction, marks grid "
                f"{len(net.enc.marks)}^3")
    base = dict(levels=4, n_feat=2, log2_T=19, n_min=r_min, per_level_scale=float(net.enc.b), num_layers=3,
                num_hidden=16, marks=net.enc.marks.numpy(), eps=1e-4, scale=1.0, describe=describe)
    rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not os.path.exists(cache) and rank != 0:
        for _ in range(3000):
            if os.path.exists(cache):
                break
            time.sleep(0.1)
    if os.path.exists(cache):
        g = np.load(cache)
        return dict(base, table=g["table"], mlp=g["mlp"])
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    sizes = net.enc.module.level_sizes
    scales, ress = [], []
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    for fn in (libm.log2f, libm.exp2f):
        fn.restype, fn.argtypes = ctypes.c_float, [ctypes.c_float]
    lp = np.float32(libm.log2f(np.float32(net.enc.b)))
    for l in range(4):
        sc = np.float32(np.float32(libm.exp2f(np.float32(l) * lp)) * np.float32(r_min) - np.float32(1))
        scales.append(float(sc)); ress.append(int(np.ceil(sc)) + 1)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    table = torch.nn.Parameter(net.enc.module.params.detach().clone().to(dev))
    fcs = [torch.nn.Linear(a, b).to(dev) for a, b in zip(net.num_nodes[:-1], net.num_nodes[1:])]
    for fc, ref in zip(fcs, net.fc):
        fc.load_state_dict(ref.state_dict())
    U32 = 0xFFFFFFFF

    def encode(x):
        tab = table.view(-1, 2)
        outs = []
        for l in range(4):
            pos = x * scales[l] + 0.5
            cf = torch.floor(pos)
            fr = pos - cf
            c = cf.long() & U32
            acc = 0
            for corner in range(8):
                w = 1
                cc = []
                for d in range(3):
                    bit = (corner >> d) & 1
                    w = w * (fr[:, d] if bit else 1 - fr[:, d])
                    cc.append((c[:, d] + bit) & U32)
                if ress[l] ** 3 <= sizes[l]:
                    idx = (cc[0] + cc[1] * ress[l] + cc[2] * ress[l] * ress[l]) & U32
                else:
                    idx = cc[0] ^ ((cc[1] * 2654435761) & U32) ^ ((cc[2] * 805459861) & U32)
                acc = acc + w.unsqueeze(-1) * tab[int(offs[l]) + idx % sizes[l]]
            outs.append(acc)
        return torch.cat(outs, -1)

    def sdf_fn(x):
        if shape == "sphere":
            return 0.6 - x.norm(dim=-1)
        q = (x[:, 0] ** 2 + x[:, 1] ** 2).sqrt() - 0.55
        return 0.22 - (q ** 2 + x[:, 2] ** 2).sqrt()

    opt = torch.optim.Adam([table] + [p for fc in fcs for p in fc.parameters()], lr=1e-2)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, steps)
    gen = torch.Generator(device=dev).manual_seed(0)
    for _ in range(steps):
        x = torch.rand(batch, 3, device=dev, generator=gen) * 2 - 1
        h = encode((x + 1) / 2)
        for i, fc in enumerate(fcs):
            h = fc(h)
            if i != len(fcs) - 1:
                h = torch.relu(h)
        loss = (torch.tanh(h[:, 1] - h[:, 0]) - sdf_fn(x).clamp(-0.3, 0.3)).abs().mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
    mlp = np.concatenate([np.concatenate([fc.weight.detach().cpu().numpy().reshape(-1),
                                          fc.bias.detach().cpu().numpy().reshape(-1)]) for fc in fcs]).astype(np.float32)
    tab = table.detach().cpu().numpy().astype(np.float32)
    tmp = cache + f".{os.getpid()}.tmp.npz"
    np.savez(tmp, table=tab, mlp=mlp, loss=float(loss.detach()))
    os.replace(tmp, cache)
    return dict(base, table=tab, mlp=mlp)


def make_native(w, pinned=None):
    from tropical._native import NativeNet
    src = pinned if pinned is not None else w
    return NativeNet(w["levels"], w["n_feat"], w["log2_T"], w["n_min"], w["per_level_scale"],
                     w["num_layers"], w["num_hidden"], src["table"], src["mlp"], src["marks"],
                     w["eps"], w["scale"])


def oracle_params(w):
    """CPU checker network -- only the cpu_baseline / --impl reference legs call this."""
    from oracle.trinet import NetParams
    nodes = [w["levels"] * w["n_feat"]] + [w["num_hidden"]] * (w["num_layers"] - 1) + [2]
    ws, bs, o = [], [], 0
    for i in range(w["num_layers"]):
        n = nodes[i] * nodes[i + 1]
        ws.append(w["mlp"][o:o + n].reshape(nodes[i + 1], nodes[i])); o += n
        bs.append(w["mlp"][o:o + nodes[i + 1]]); o += nodes[i + 1]
    return NetParams(w["levels"], w["n_feat"], w["log2_T"], w["n_min"], w["per_level_scale"],
                     w["num_layers"], w["num_hidden"], w["table"], ws, bs, w["marks"], w["eps"], w["scale"])


# ---------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  Uses NVML in-process
    (a sampling thread); an external `nvidia-smi -lms` process was measured to stall CUDA
    calls of the benchmarked process for tens of milliseconds while it initialises."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, period=0.02):
        self.idx, self.period, self.rows, self.run, self.thread, self.nv = gpu_index, period, [], False, None, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.run = True
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if self.idx < len(ids) and ids[self.idx].strip().isdigit():
                return int(ids[self.idx])
        return self.idx

    def mark(self):
        """Samples taken from now on belong to the timed region."""
        self.rows = []

    def _pump(self):
        nv = self.nv
        while self.run:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((mhz, mask))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self.run = False
        if self.thread:
            self.thread.join(timeout=1.0)
        sm = [r[0] for r in self.rows]
        reasons = sorted({name for _, mask in self.rows for bit, name in self.REASONS.items() if mask & bit})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nv else "unavailable"}


# ---------------------------------------------------------------------------------------
# the CPU arm (oracle port of the reference's algorithm)
# ---------------------------------------------------------------------------------------
def cpu_extract_seconds(P, planar=True):
    from oracle import subpoly_ref as R
    t = time.perf_counter()
    faces, vertices, tri = R.subpoly(P, force=planar)
    return time.perf_counter() - t, vertices.shape[0], tri.shape[0]


_WL_KEYS = ("levels", "n_feat", "log2_T", "n_min", "per_level_scale", "num_layers", "num_hidden", "table", "mlp", "marks", "eps", "scale")


def reference_worker_main(path, planar, n_mine, warm):
    """`bench.py --ref-worker ...`: one host core's share of the run's full extractions with the oracle port, in a
    process of its own that imports numpy and the C checker only (no torch, no CUDA: the parent fitted the workload
    on the GPU, and a forked copy of that process hung in the OpenMP / CUDA state it inherited)."""
    g = np.load(path)
    w = {k: (g[k] if g[k].ndim else g[k].item()) for k in _WL_KEYS}
    P = oracle_params(w)
    if warm:   # page in numpy / the C checker with a small extraction (untimed)
        cpu_extract_seconds(oracle_params(load_workload("small_sphere")), planar)
    t0 = time.perf_counter()
    nv = nt = 0
    for _ in range(n_mine):
        _, nv, nt = cpu_extract_seconds(P, planar)
    print(json.dumps({"seconds": time.perf_counter() - t0, "nv": int(nv), "nt": int(nt)}), flush=True)


def shared_config(w, planar, nv, nt):
    """`config` of the JSON line: identical in both arms (the mesh sizes are bit-exact parity facts)."""
    return {"workload": w["describe"] + ("" if planar else " [curve-approximation path]"),
            "path": "planar" if planar else "curve", "marks_grid": int(len(w["marks"])),
            "mesh_vertices": int(nv), "mesh_triangles": int(nt)}


def run_reference(args):
    """The reference's algorithm on the box's host cores (oracle port; the reference itself is Python
    over tiny-cuda-nn and cannot run on the GPU box).  The path shards by object: the run's K full
    extractions are dealt out over the worker processes, one per core, and
    value = vertices of all K extractions / wall time of the slowest worker."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tempfile
    w = load_workload(args.workload)
    planar = args.path == "planar"
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # bounded so that the workers' working sets (numpy temporaries of [sum 2^k, 36] int64 region matrices:
    # gigabytes each for the large model) cannot exhaust the box's memory
    cap = args.ref_cores if args.ref_cores > 0 else {"small": 32, "medium": 16, "large": 8}.get(args.workload.split("_")[0], 8)
    cores = max(1, min(avail, cap, args.steps))
    share = [args.steps // cores + (1 if i < args.steps % cores else 0) for i in range(cores)]
    fd, path = tempfile.mkstemp(suffix=".npz", prefix="tnb_ref_workload_")
    os.close(fd)
    np.savez(path, **{k: np.asarray(w[k]) for k in _WL_KEYS})
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--ref-worker", path, str(int(planar)), str(n),
                               str(1 if args.warmup > 0 else 0)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
             for n in share]
    res = []
    for pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            raise SystemExit("reference worker failed: " + err[-2000:])
        r = json.loads(out.strip().splitlines()[-1])
        res.append((r["seconds"], r["nv"], r["nt"]))
    os.unlink(path)
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    nv, nt = res[0][1], res[0][2]
    value = nv * args.steps / slowest
    cpu = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{args.steps} full extraction(s) of the same network dealt out over {cores} worker process(es), one per host "
                     f"core ({max(share)} each at most), numpy+C oracle port; warm-up = one small-model extraction per worker",
           "wall_s": wall, "seconds_per_extraction_one_core": slowest / max(share)}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * slowest / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": shared_config(w, planar, nv, nt),
                      "cpu_baseline": cpu,
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from tropical import _native
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = load_workload(args.workload)
    net = make_native(w)
    R = net.n_outputs
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    planar = args.path == "planar"
    slab = args.shard == "slab" and world > 1
    sweep_sharded = args.shard == "sweep" and world > 1
    if slab and not planar:
        raise SystemExit("--shard slab supports the planar path only")

    class SlabMesh:
        """One object sharded over the ranks by marks-grid slabs and merged (tropical/parallel.py)."""

        def __init__(self, n, phase_ms=False):
            from tropical import parallel
            self.v, self.t, self.stats = parallel.subpoly_sharded(n, size=1.2, eps=w["eps"], phase_ms=phase_ms)

        def sizes(self):
            return {"V": int(self.v.shape[0]), "T": int(self.t.shape[0]), "P": None}

        def read_host(self, polygons=False):
            v, t = self.v.cpu().numpy(), self.t.cpu().numpy()
            return v, t, v[t], None

    def step(n=None):
        if slab:
            return SlabMesh(n or net)
        if sweep_sharded:   # ONE object: the skeleton sweep dealt to the ranks by grid planes, the rest on every rank
            from tropical import parallel
            return parallel.subpoly_sweep_sharded(n or net, size=1.2, eps=w["eps"], force=planar)
        return (n or net).subpoly(size=1.2, eps=w["eps"], force=planar)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        mesh = step()
    sizes = mesh.sizes()
    slab_stats = dict(mesh.stats) if slab else None
    del mesh

    _native.profile_enable(True)
    _native.profile_reset()
    _native.lib().tnb_launch_count_reset()
    barrier()
    sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.fill_(1)          # L2 flush between timed iterations (outside the event pair)
        a.record()
        mesh = step()
        b.record()
        del mesh
    barrier()
    launches = int(_native.lib().tnb_launch_count())
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    prof = _native.profile_read()
    _native.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    rank_ms = None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # every rank's mean / fastest / slowest step: the headline takes the slowest rank's sum
        mine = torch.tensor([sum(step_ms) / len(step_ms), min(step_ms), max(step_ms)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_ms = [[round(float(x), 4) for x in r.tolist()] for r in allr]
    ms_total = float(t.item())
    objects = 1 if (slab or sweep_sharded) else world   # slab / sweep mode: ONE object over all ranks (strong scaling)
    value = sizes["V"] * objects * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C ABI with host buffers --------------------------------
    pinned = {k: torch.from_numpy(np.ascontiguousarray(w[k], np.float32)).pin_memory() for k in ("table", "mlp", "marks")}
    pinned_np = {k: v.numpy() for k, v in pinned.items()}
    h2d = sum(v.numel() * 4 for v in pinned.values())
    d2h = 0

    # pinned result buffers, reused from step to step (a fresh pageable array per step costs its page faults: ~1 ms for
    # the 19 MB of the large mesh); sized with head-room over the warm-up's mesh
    res = None
    if not slab:
        cap_v, cap_t = int(sizes["V"] * 1.25) + 1024, int(sizes["T"] * 1.25) + 1024
        res_t = {"v": torch.empty(cap_v * 3, dtype=torch.float32).pin_memory(), "t": torch.empty(cap_t * 3, dtype=torch.int64).pin_memory(),
                 "f": torch.empty(cap_t * 9, dtype=torch.float32).pin_memory()}
        res = {k: x.numpy() for k, x in res_t.items()}

    def e2e_step():
        n2 = make_native(w, pinned_np)            # host -> device copy of the step's inputs
        m2 = step(n2)
        if res is not None and m2.sizes()["V"] <= cap_v and m2.sizes()["T"] <= cap_t:
            v, tr, f, _ = m2.read_host(polygons=False, out=res)   # device -> host read of subpoly()'s return values
        else:
            v, tr, f, _ = m2.read_host(polygons=False)
        return v.nbytes + tr.nbytes + f.nbytes

    for _ in range(2):
        d2h = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = sizes["V"] * objects * args.steps / float(t.item())

    # ---- strong scaling of ONE object over the N GPUs (outside the timed region; every rank takes part) ------
    strong = None
    if world > 1 and not slab and not sweep_sharded and args.strong != "none":
        from tropical import parallel
        strong = {"note": "ONE object over all ranks (the headline above is one object per rank); ms = max over ranks, CUDA events, "
                          "L2 flushed between steps; mesh compared with this rank's single-GPU mesh"}
        ref_sizes = sizes

        def timed(fn, reps):
            for _ in range(3):
                m = fn()
            got = m.sizes() if hasattr(m, "sizes") else None
            barrier()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            for a, b in evs:
                flush.fill_(1)
                a.record()
                m = fn()
                b.record()
                del m
            barrier()
            tt = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / reps], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()), got

        if "sweep" in args.strong:
            ms_sw, got = timed(lambda: parallel.subpoly_sweep_sharded(net, size=1.2, eps=w["eps"], force=planar), args.steps)
            strong["sweep_sharded"] = {"ms_per_extraction": ms_sw, "vertices_per_s": sizes["V"] / (ms_sw * 1e-3),
                                       "speedup_vs_one_gpu": (ms_total / args.steps) / ms_sw,
                                       "mesh_equals_single_gpu": bool(got and got["V"] == ref_sizes["V"] and got["T"] == ref_sizes["T"]),
                                       "what": "skeleton sweep dealt to the ranks by planes of the marks grid; one NCCL all-gather of the |sdf| "
                                               "planes + one MAX all-reduce of the per-chunk gradient maxima; subdivision and faces on every rank: "
                                               "bit-identical to the single-GPU mesh by construction"}
        if "slab" in args.strong and planar:
            try:
                ms_sl, _ = timed(lambda: SlabMesh(net), max(2, args.steps // 2))
                sm = SlabMesh(net, phase_ms=True)
                strong["slab_sharded"] = {"ms_per_extraction": ms_sl, "speedup_vs_one_gpu": (ms_total / args.steps) / ms_sl,
                                          "phase_ms_rank0": sm.stats.get("phase_ms"),
                                          "merged_vertices": sm.sizes()["V"], "merged_triangles": sm.sizes()["T"],
                                          "single_gpu_vertices": ref_sizes["V"], "single_gpu_triangles": ref_sizes["T"],
                                          "near_plane": sm.stats.get("near_plane"),
                                          "what": "cell slabs along the first axis, per-hyperplane liveness exchange through peer mailboxes (NVLink), "
                                                  "all-gather + merge; exact iff near_plane == 0"}
            except Exception as e:  # noqa: BLE001  (a failed optional leg must not take the headline down)
                strong["slab_sharded"] = {"error": str(e)[:300]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel class -------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    top = max(prof, key=lambda k: prof[k][0])
    ms, n_launch, units, algo = prof[top]   # algorithmic bytes are accounted by the library
    achieved = algo / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    tj_all = json.load(open(tpath)) if os.path.exists(tpath) else {}
    tj = tj_all.get(args.workload, {}).get(top)
    if tj:
        traffic = tj["dram_bytes_per_launch"]
    # what actually bounds each kernel class (ncu --set full, profiles/r2_ncu_summary.txt); `hbm` is the contract's word
    # for the denominator, not a claim that these kernels are bandwidth bound
    bound_notes = {
        "sweep": "fp32 issue / instruction fetch: 1 136 instructions per grid vertex (sdf + input gradient), issue slots 55 % busy, FMA pipe 26 %, "
                 "stall no_instruction 0.95 warps per issue (113 KB of straight-line code), 168 registers -> 12 warps per SM; HBM traffic is the |sdf| array only",
        "vertex_rows": "fp32 issue + L2 gathers; rows leave through shared memory with coalesced stores",
        "step_front": "new vertices: L2 latency (two dependent gathers per crossed edge, then a network evaluation by 1 thread in ~30)",
        "step_back": "connecting edges: L2 latency (cell header -> segment records, long_scoreboard 4.1 warps per issue) and shuffle traffic of the partner sort",
        "face_rows": "face rows: L2 latency (one warp per surface vertex streams the sorted segment of the region's cell)",
        "sign_sweep": "fp32 issue: 1 002 instructions per lattice point, issue slots 62 % busy, FMA pipe 29 %",
        "steps_persistent": "grid-barrier / dependent-load latency: 7 barriers (~2.5 us each) per crossing hyperplane, issue slots 6 % busy; the complex stays in L2",
    }
    # fp32 view of the evaluation kernels: FMA = 2 flop; a grid vertex of the skeleton sweep costs the forward pass
    # (64 interpolation + 416 MLP FMAs, 48 weight products) and the input gradient (416 MLP + 96 interpolation FMAs, 96 differences)
    clock_ghz = (clocks or {}).get("sm_mhz", 1965.0) / 1e3 if clocks else 1.965
    fp32_peak = 148 * 128 * 2 * clock_ghz / 1e3   # TFLOP/s at the clock sampled under load
    flops_per_unit = {"sweep": 2 * (64 + 416 + 416 + 96) + 48 + 96, "vertex_rows": 2 * (64 + 416) + 48, "sign_sweep": 2 * (64 + 416) + 48}
    by_kernel = {}
    for k, (kms, kn, ku, kb) in prof.items():
        if kn == 0:
            continue
        one = {"ms_per_step": kms / args.steps, "launches_per_step": kn / args.steps, "units_per_step": ku / args.steps,
               "algorithmic_bytes_per_step": kb / args.steps, "achieved_gbs": kb / (kms * 1e-3) / 1e9 if kms > 0 else 0.0,
               "hbm_frac": (kb / (kms * 1e-3) / 1e9 / peak) if kms > 0 else 0.0, "share_of_step": kms / ms_total,
               "what_bounds_it": bound_notes.get(k, "")}
        if k in flops_per_unit and kms > 0:
            tf = ku * flops_per_unit[k] / (kms * 1e-3) / 1e12
            one["fp32"] = {"achieved_tflops": tf, "peak_tflops": fp32_peak, "frac": tf / fp32_peak, "flops_per_unit": flops_per_unit[k]}
        kt = tj_all.get(args.workload, {}).get(k)
        if kt:
            one["dram_bytes_per_launch_ncu"] = kt["dram_bytes_per_launch"]
        by_kernel[k] = one
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo / max(n_launch, 1), "launches": n_launch,
                "avg_launch_ms": ms / max(n_launch, 1), "kernel_share_of_step": ms / ms_total,
                "units": units, "units_per_launch": units / max(n_launch, 1),
                "algorithmic_bytes_per_unit": algo / max(units, 1),
                "note": "algorithmic bytes = compulsory first-touch traffic of the class (outputs once, gathered rows once, tables once per launch; "
                        "DESIGN.md section 5). The dominant class is NOT bandwidth bound: " + bound_notes.get(top, "") + ". Timed classes cover "
                        "%.0f %% of the step; the rest are the ordered compactions and housekeeping kernels between them." %
                        (100 * sum(v[0] for v in prof.values()) / ms_total),
                # the fp32 view of the dominant class if it evaluates the network, else of the sweep (k_sweep_pieces: the largest
                # single launch of a large extraction, and the one kernel here whose ceiling is arithmetic)
                "fp32": by_kernel.get(top, {}).get("fp32") or
                        (dict(by_kernel["sweep"]["fp32"], kernel="sweep", share_of_step=by_kernel["sweep"]["share_of_step"])
                         if by_kernel.get("sweep", {}).get("fp32") else None),
                "by_kernel": by_kernel,
                "by_kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
                "by_kernel_gbs": {k: (v[3] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0) for k, v in prof.items()}}

    # ---- evaluation-sweep throughput (BASELINE configs[4]), outside the timed region -----
    sweep = None
    if not args.no_sweep:
        sweep = {"algorithmic_bytes_per_point": 16,
                 "note": "batched trilinear network eval + packed sign vectors over a dense lattice; output larger than L2 "
                         "(16 B/point); the kernel is fp32-issue bound", "lattices": []}
        for n3 in ([args.sweep_n] if args.sweep_n > 0 else [256, 512, 1024]):
            buf = torch.empty((n3 ** 3, 2), dtype=torch.int64, device="cuda")
            for _ in range(2):
                net.sweep_signs((-1, -1, -1), (1, 1, 1), (n3, n3, n3), out=buf)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                net.sweep_signs((-1, -1, -1), (1, 1, 1), (n3, n3, n3), out=buf)
            b.record()
            torch.cuda.synchronize()
            ms_sw = a.elapsed_time(b) / reps
            pts = n3 ** 3
            one = {"lattice": f"{n3}^3", "points_per_s": pts / (ms_sw * 1e-3), "ms": ms_sw,
                   "achieved_gbs": pts * 16 / (ms_sw * 1e-3) / 1e9, "hbm_frac": pts * 16 / (ms_sw * 1e-3) / 1e9 / peak,
                   "approx_fp32_tflops": pts * 2 * (4 * 8 * 5 + 8 * 16 + 16 * 16 + 16 * 2) / (ms_sw * 1e-3) / 1e12}
            sweep["lattices"].append(one)
            if n3 == 512 or len(sweep["lattices"]) == 1:
                sweep.update(one)
            del buf
            torch.cuda.empty_cache()

    # ---- several objects in flight on ONE GPU (outside the timed region; N=1 only) ---------
    # One call of tnb_subpoly_batch: K worker threads of the library, one CUDA stream each; the hyperplanes of
    # every small object run inside ONE thread-block cluster (16 SMs), so their step loops are resident side by side.
    concurrent = None
    if world == 1 and not slab and planar and args.concurrent > 1:
        K, M = args.concurrent, 12
        # the SMALL model (BASELINE configs[1]): objects whose hyperplanes fit one 16-CTA cluster; the headline workload
        # itself follows below when it is another one
        ws = w if args.workload.startswith("small") else load_workload("small_sphere")
        net_s = net if ws is w else make_native(ws)

        def step_s():
            return net_s.subpoly(size=1.2, eps=ws["eps"], force=True)

        for _ in range(3):
            ms_ = step_s()
        v_small = ms_.sizes()["V"]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(K * 2):
            step_s()
        torch.cuda.synchronize()
        one_at_a_time = K * 2 * v_small / (time.perf_counter() - t1)
        # one call for a batch of objects (tnb_subpoly_batch): K of them in flight on K worker streams of the library
        B = 4 * K
        for _ in range(2):
            ms_ = _native.subpoly_batch([net_s] * B, size=1.2, eps=ws["eps"], force=True, in_flight=K)
        assert all(m_.sizes()["V"] == v_small for m_ in ms_)
        del ms_
        torch.cuda.synchronize()
        tc = time.perf_counter()
        for _ in range(M):
            ms_ = _native.subpoly_batch([net_s] * B, size=1.2, eps=ws["eps"], force=True, in_flight=K)
            del ms_
        torch.cuda.synchronize()
        dtc = time.perf_counter() - tc
        concurrent = {"workload": ws["describe"], "entry": "tnb_subpoly_batch", "objects_per_call": B, "objects_in_flight": K,
                      "objects_per_s": B * M / dtc, "vertices_per_s": B * M * v_small / dtc,
                      "one_at_a_time_vertices_per_s": one_at_a_time, "vs_one_at_a_time": (B * M * v_small / dtc) / one_at_a_time,
                      "note": "one call, K worker threads x 1 stream inside the library; small complexes run their hyperplanes in one "
                              "16-CTA cluster each (k_steps_cluster); work buffers come from the library's block cache (the stream-ordered "
                              "allocator calls were what serialised the host threads before)"}

        if ws is not w:
            # the headline workload, 4 in flight: a large extraction is a stream of ~170 launches, many of them far
            # from filling 148 SMs, so a second and a third object find room next to it
            Kb, Bb, Mb = 4, 8, 3
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(Bb):
                step()
            torch.cuda.synchronize()
            one_big = Bb / (time.perf_counter() - t1)
            ms_ = _native.subpoly_batch([net] * Bb, size=1.2, eps=w["eps"], force=True, in_flight=Kb)
            assert all(m_.sizes()["V"] == sizes["V"] for m_ in ms_)
            del ms_
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(Mb):
                ms_ = _native.subpoly_batch([net] * Bb, size=1.2, eps=w["eps"], force=True, in_flight=Kb)
                del ms_
            torch.cuda.synchronize()
            per_s = Bb * Mb / (time.perf_counter() - t1)
            concurrent["headline_workload"] = {"entry": "tnb_subpoly_batch", "objects_per_call": Bb, "objects_in_flight": Kb,
                                               "objects_per_s": per_s, "vertices_per_s": per_s * sizes["V"],
                                               "one_at_a_time_objects_per_s": one_big, "vs_one_at_a_time": per_s / one_big,
                                               "note": "back to back, no L2 flush between objects (both figures)"}

    # ---- CPU baseline (bounded sample on the box's host cores) ----------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        dt, nv, _ = cpu_extract_seconds(oracle_params(w), planar)
        cpu = {"value": nv / dt, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
               "sample": "1 full extraction of the same network with the numpy+C oracle port"}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
           "scaling": "strong" if (slab or sweep_sharded) else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": shared_config(w, planar, sizes["V"], sizes["T"]),
           "run": {"polygons": sizes["P"], "objects_per_step": objects, "rank_ms_mean_min_max": rank_ms, "l2": "flushed (512 MiB write) between timed steps",
                   "extraction_s": ms_total / args.steps * 1e-3,
                   "sharding": ("one object cut into %d marks-grid slabs, one per GPU; per-step exchange through peer mailboxes, "
                                "all-gather + merge inside the timed region" % world) if slab else
                               ("one object; skeleton sweep dealt to %d GPUs by grid planes (all-gather inside the timed region), the rest on every rank" % world)
                               if sweep_sharded else "one object per GPU",
                   "slab_stats": slab_stats},
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": 1e3 * float(t.item()) / args.steps},
           "roofline": roofline, "cpu_baseline": cpu, "eval_sweep": sweep, "concurrent": concurrent, "strong_scaling": strong}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    if len(sys.argv) >= 6 and sys.argv[1] == "--ref-worker":
        return reference_worker_main(sys.argv[2], sys.argv[3] == "1", int(sys.argv[4]), sys.argv[5] == "1")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="large_sphere",
                    help="large_sphere (default: the configuration BASELINE.json quotes its target on), medium_sphere, "
                         "small_sphere, small_torus, medium_torus, large_torus, <size>_random")
    ap.add_argument("--path", default="planar", choices=["planar", "curve"],
                    help="planar = the reference's -f default (force=True); curve = curve approximation")
    ap.add_argument("--shard", default="object", choices=["object", "slab", "sweep"],
                    help="N>1: object = every rank extracts its own object (weak scaling, default); "
                         "slab = ONE object cut into marks-grid slabs, one per GPU (strong scaling); "
                         "sweep = ONE object, its skeleton sweep dealt to the GPUs by grid planes (strong scaling, exact)")
    ap.add_argument("--strong", default="sweep", help="N>1, --shard object: extra strong-scaling legs of ONE object, outside the timed region: "
                                                      "'sweep' (default), 'sweep,slab', or 'none'")
    ap.add_argument("--ref-cores", type=int, default=0, help="--impl reference: host cores to use (0 = all)")
    ap.add_argument("--concurrent", type=int, default=8, help="extra leg: objects in flight on one GPU (0/1 = skip)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the evaluation-sweep throughput leg")
    ap.add_argument("--sweep-n", type=int, default=0, help="lattice size per axis of the evaluation sweep (0 = 256, 512 and 1024)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
