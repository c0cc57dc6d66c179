#!/usr/bin/env python
"""bench.py -- throughput of the ParticleEmitter hot path (update + billboard draw per frame) on B200.

    python bench.py --gpus N --steps K --warmup W            the CUDA backend through the C ABI
    python bench.py --impl reference --gpus N ...            the reference's own CPU path on the host cores

A step is one frame over the whole population: `update(1000/60 ms)` then `draw()` for every emitter
(PE.cpp:713-888).  Workloads (BASELINE.json configs / SURVEY.md §8(d)):
    c3  one emitter, 64 Mi particles, ellipsoid spawn, per-particle spin, looped sprite animation (N = 1 default)
    c4  8 192 emitters x 32 Ki = 256 Mi particles, mixed fire/smoke/explosion sets, emitter e owned by rank e mod N
        (N > 1 default; strong scaling, no per-frame collective; NCCL only gathers the end-of-run counters)
    c2  1 024 emitters x 4 Ki (4 Mi particles), mixed sets
    c5  high churn: 8 emitters per GPU, 1 Mi spawned per frame PER GPU, energies 50..400 ms (7 % of the live set replaced
        per frame: spawn- and compaction-bound)
    c3d the c3 emitter with deaths: 1 Mi spawned per frame, energies 500..1500 ms (1.7 % replaced per frame, ~63 M live)
The N = 1 line also carries short measurements of c2, c3d, c4 and c5 under `also` (every BASELINE.json config on one
GPU; `also.c4` is the N = 1 point of the strong-scaling curve the N > 1 lines continue, `also.c5` that of the churn
workload, which is weak-scaled: 1 Mi spawned per frame per GPU).
One JSON line is printed by rank 0.  `value` = particles x steps / device time of the timed region with all
state resident in HBM; `e2e` = the same frames driven call by call through the public ParticleEmitter
interface with the per-frame host inputs (node and camera matrices, elapsed time) coming from host memory and
the per-frame results (live count, quad count) read back to the host every frame.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DT = 1000.0 / 60.0
MI = 1024 * 1024

# Algorithmic bytes per unit, SURVEY.md Appendix C (materialised colour/size layout actually shipped):
#   update: energy rw 8 + energyStart 4 + pos rw 24 + vel rw 24 + acc 12 + spin 4 + angle rw 8
#           + colorStart/End 32 + color w 16 + sizeStart/End 8 + size w 4            = 144 B
#           + frame rw 8 + timeOnFrame rw 8 when the sprite is animated               = 160 B
#   draw:   pos 12 + size 4 + color 16 + angle 4 + frame 4 read, 4 x 36 B written     = 184 B
#   spawn:  34 columns x 4 B written                                                   = 136 B
UPDATE_BYTES, UPDATE_BYTES_ANIMATED, DRAW_BYTES, SPAWN_BYTES = 144, 160, 184, 136


def measured_traffic():
    """DRAM bytes per unit of each kernel from the committed ncu capture (profiles/), or {}."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode()) if uuid else pynvml.nvmlDeviceGetHandleByIndex(index)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is None:
            return
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------
# churn knobs (particles spawned per frame PER GPU, energy range in ms); --churn-spawn / --churn-energy change c5's
CHURN = {"c5": {"spawn": MI, "energy": (50.0, 400.0), "emitters": 8, "warm": 30},
         "c3d": {"spawn": MI, "energy": (500.0, 1500.0), "emitters": 1, "warm": 100}}


def build_population(ctx, workload, rank, world, particles):
    """Creates this rank's emitters and fills them.  Returns (emitters, per_emitter_particles, animated_fraction, name)."""
    from tractor3d_b200 import abi, presets, sharding
    from tractor3d_b200.emitter import ParticleEmitter

    emitters = []
    if workload == "c3":
        n = particles or 64 * MI
        p, sprite = presets.headline_64m(capacity=n)
        e = ParticleEmitter.create(ctx, p, sprite, (abi.RNG_DEVICE, 1234 + rank))
        e.emitOnce(n)
        emitters.append(e)
        name = f"c3: 1 emitter x {n} particles, ellipsoid spawn, per-particle spin, looped 4-frame sprite animation"
        return emitters, n, 1.0, name
    if workload in ("c4", "c2"):
        n_emitters, per = (8192, 32 * 1024) if workload == "c4" else (1024, 4096)
        if particles:
            per = max(32, particles // n_emitters)
        animated = 0
        for eid in sharding.local_emitter_ids(n_emitters, rank, world):
            p, sprite = presets.MIXED[eid % 3]()
            p.particle_count_max = per
            p.energy_min, p.energy_max = 1.0e6, 2.0e6   # stationary population
            e = ParticleEmitter.create(ctx, p, sprite, (abi.RNG_DEVICE, eid))
            emitters.append(e)
            animated += 1 if p.sprite_animated else 0
        ctx.batch_emit_once(emitters, [per] * len(emitters))
        name = (f"{workload}: {n_emitters} emitters x {per} particles, fire/smoke/explosion parameter sets round-robin, "
                f"emitter e owned by rank e mod {world}")
        return emitters, per, animated / max(1, len(emitters)), name
    if workload in CHURN:
        k = CHURN[workload]
        n_em, cap = k["emitters"], (particles or 64 * MI) // k["emitters"]
        for j in range(n_em):
            p, sprite = presets.headline_64m(capacity=cap)
            p.energy_min, p.energy_max = k["energy"]
            e = ParticleEmitter.create(ctx, p, sprite, (abi.RNG_DEVICE, 77 + rank * n_em + j))
            emitters.append(e)
        name = (f"{workload}: {n_em} emitter(s)/GPU, {k['spawn']} particles spawned per frame per GPU, "
                f"energies {k['energy'][0]:g}..{k['energy'][1]:g} ms")
        return emitters, cap, 1.0, name
    raise SystemExit(f"unknown workload {workload}")


def kernel_rooflines(kern, particles_per_launch, animated_frac, moved_per_launch=0, spawned_per_launch=0, traffic=None):
    """Per-kernel roofline objects from average launch times (CUDA events on the launching stream)."""
    peak, peak_src = measured_peak_gbs()
    upd_b = UPDATE_BYTES + (UPDATE_BYTES_ANIMATED - UPDATE_BYTES) * animated_frac
    # fused: the draw's 40 B re-read of position/colour/size/angle/frame disappears (the frame index is read once)
    fused_b = upd_b + DRAW_BYTES - 40 + (0 if animated_frac == 1.0 else 4 * (1 - animated_frac))
    # compaction: every particle that dies pulls one in from the end of the array, 2 x 136 B
    move_bytes = 2 * SPAWN_BYTES * moved_per_launch
    roof = {}
    for nm, b, extra, units in (("update", upd_b, move_bytes, particles_per_launch), ("draw", DRAW_BYTES, 0, particles_per_launch),
                                ("update_draw", fused_b, move_bytes, particles_per_launch), ("spawn", SPAWN_BYTES, 0, spawned_per_launch)):
        ms = kern.get(nm, {}).get("ms_per_launch")
        if not ms or not units:
            continue
        ach = (b * units + extra) / (ms * 1e-3) / 1e9
        tr = (traffic or {}).get(nm, {}).get("bytes_per_unit")
        roof[nm] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": tr * units if tr else None, "bytes_per_unit": b, "units_per_launch": units, "ms_per_launch": ms,
                    "peak_source": peak_src}
        if extra:
            roof[nm]["moved_per_launch"] = moved_per_launch
    return roof


def kernel_deltas(t0, t1):
    """{kernel: ms per launch, launches, total ms} between two readings of the context's per-kind event timers."""
    out = {}
    for k, nm in enumerate(("spawn", "update", "draw", "update_draw")):
        dms, dn = t1[k][0] - t0[k][0], t1[k][1] - t0[k][1]
        if dn:
            out[nm] = {"ms_per_launch": dms / dn, "launches": int(dn), "ms_total": dms}
    return out


def quick_measure(ctx, stream, workload, steps, torch, rank=0, world=1):
    """A short extra measurement of another workload on the same context (reported under `also`)."""
    emitters, per, animated_frac, wl_name = build_population(ctx, workload, rank, world, 0)
    handles = ctx._handles(emitters)
    churn = CHURN.get(workload)
    spawn_each = churn["spawn"] // len(emitters) if churn else 0

    def frame():
        if churn:
            ctx.batch_emit_once(emitters, [spawn_each] * len(emitters), handles)
        ctx.batch_update(emitters, DT, handles)
        ctx.batch_draw(emitters, handles)

    for _ in range(churn["warm"] if churn else 5):
        frame()
    ctx.sync()
    n0 = int(ctx.batch_counts(emitters, handles).sum())
    t0 = [ctx.kernel_time(k) for k in range(4)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            frame()
        ctx.flush()
        ev1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    t1 = [ctx.kernel_time(k) for k in range(4)]
    n1 = int(ctx.batch_counts(emitters, handles).sum())
    n = (n0 + n1) // 2
    kern = kernel_deltas(t0, t1)
    moved = spawn_each * len(emitters)   # steady state: deaths per frame = spawns per frame
    out = {"workload": wl_name, "particles": n, "steps": steps, "ms_per_step": ms / steps, "value": n * steps / (ms * 1e-3),
           "unit": "particles/s", "kernels": kern,
           "rooflines": kernel_rooflines(kern, n, animated_frac, moved, moved, measured_traffic().get(workload))}
    if churn:
        out["replaced_per_frame"] = moved / max(1, n)
    for e in emitters:
        e.release()
    return out


def vertices_to_host(ctx, e, n, torch, reps=3):
    """Frames whose vertex stream is delivered to pinned host memory: drawn into one of two device buffers
    (t3d_emitter_draw_to), copied out on the context's download stream while the next frame is computed."""
    quad_bytes = 144
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
    except Exception:
        avail = 0
    n_host = 2 if avail > 6 * n * quad_bytes else 1
    try:
        dev = [torch.empty(n * 36, dtype=torch.float32, device="cuda") for _ in range(2)]
        host = [torch.empty(n * 36, dtype=torch.float32, pin_memory=True) for _ in range(n_host)]
    except Exception as err:   # not enough pinnable memory on this box: say so instead of failing the run
        return {"unavailable": f"could not allocate the buffers: {err}"}
    nbytes = n * quad_bytes
    # PCIe roofline of this box: one plain copy of the same size
    ctx.sync()
    t0 = time.perf_counter()
    ctx.copy_wait(ctx.copy_to_host_async(host[0].data_ptr(), dev[0].data_ptr(), nbytes))
    pcie_gbs = nbytes / (time.perf_counter() - t0) / 1e9
    tickets = [None, None]
    t0 = time.perf_counter()
    for k in range(reps):
        b = k & 1
        if tickets[b] is not None:
            ctx.copy_wait(tickets[b])      # buffer b's previous frame has left the device: it may be redrawn
        if n_host == 1 and tickets[b ^ 1] is not None:
            ctx.copy_wait(tickets[b ^ 1])  # (one host buffer only: the previous frame must have landed before it is overwritten)
            tickets[b ^ 1] = None
        e.update(DT)
        e.draw_to(dev[b].data_ptr(), n)
        tickets[b] = ctx.copy_to_host_async(host[b % n_host].data_ptr(), dev[b].data_ptr(), nbytes)
    for t in tickets:
        if t is not None:
            ctx.copy_wait(t)
    dt = time.perf_counter() - t0
    out = {"value": n * reps / dt, "unit": "particles/s", "d2h_bytes_per_step": nbytes, "steps": reps,
           "pcie_d2h_gbs_measured": pcie_gbs, "pcie_roofline_particles_per_s": pcie_gbs * 1e9 / quad_bytes,
           "frac_of_pcie": (n * reps / dt) / (pcie_gbs * 1e9 / quad_bytes), "host_buffers": n_host,
           "what": "update + draw_to(device buffer k mod 2) + asynchronous copy of the 144 B/quad vertex stream to pinned host "
                   "memory, overlapped with the next frame; bound by the PCIe link, not by the kernels"}
    del dev, host
    torch.cuda.empty_cache()
    return out


def run_cuda(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if rank == 0:
            print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun --nproc-per-node {args.gpus}", file=sys.stderr)
        if world == 1 and args.gpus > 1:
            raise SystemExit(2)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the particle backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from tractor3d_b200 import abi
    from tractor3d_b200.emitter import Context

    workload = args.workload or ("c3" if world == 1 else "c4")
    stream = torch.cuda.Stream()
    ctx = Context(local_rank, stream.cuda_stream)
    if args.rotation == "per-particle":
        ctx.set_rotation_mode(abi.ROT_PER_PARTICLE)
    # T3D_BENCH_EMULATE_WORLD=N (experiments only): rank 0's share of an N-GPU run on one GPU
    pop_world = int(os.environ.get("T3D_BENCH_EMULATE_WORLD", world)) if world == 1 else world
    emitters, per, animated_frac, wl_name = build_population(ctx, workload, rank, pop_world, args.particles)
    handles = Context._handles(emitters)
    churn = workload in CHURN
    spawn_per_frame = CHURN[workload]["spawn"] // len(emitters) if churn else 0

    def frame():
        if churn:
            ctx.batch_emit_once(emitters, [spawn_per_frame] * len(emitters), handles)
        ctx.batch_update(emitters, DT, handles)
        ctx.batch_draw(emitters, handles)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # churn needs ~25 frames (the longest lifetime) before the live population is stationary
    n_warm = max(CHURN[workload]["warm"] if churn else 3, args.warmup)
    for _ in range(n_warm):
        frame()
    barrier()
    counts0 = ctx.batch_counts(emitters, handles)   # live population at the start of the timed region
    local_particles = int(counts0.sum())

    # ---- timed region: K frames back to back, device time on the launching stream ----
    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(local_rank), "uuid", None))
    sampler.start()
    t_before = [ctx.kernel_time(k) for k in range(4)]
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            frame()
        ctx.flush()
        ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.result()
    t_after = [ctx.kernel_time(k) for k in range(4)]
    launches = ctx.launch_count() - launches0
    counts1 = ctx.batch_counts(emitters, handles)
    live_after = int(counts1.sum())
    updates_done = (local_particles if not churn else int((counts0.sum() + counts1.sum()) // 2)) * args.steps
    local_updates_done = updates_done   # this rank's share: the per-launch rooflines below are rank 0's kernels

    kern = kernel_deltas(t_before, t_after)

    # ---- e2e: the same frames through the per-emitter public calls, host inputs in, host results out, every frame ----
    import numpy as np
    e2e_steps = max(3, min(args.steps, 20))
    nodes = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (len(emitters), 1))
    cam = np.eye(4, dtype=np.float32).reshape(16)
    single = len(emitters) == 1
    e2e_counts = np.zeros(len(emitters), dtype=np.uint32)

    def e2e_run(pipelined):
        barrier()
        b0 = ctx.transfer_bytes()
        t0 = time.perf_counter()
        units = 0
        for s in range(e2e_steps):
            nodes[:, 12] = 0.001 * s
            if single:
                # exactly the calls a game makes on one emitter (ParticlesSample.cpp:897-898, 916-921)
                e = emitters[0]
                e.set_node_world(nodes[0])  # host -> library, every frame (Node::getWorldMatrix, PE.cpp:299)
                e.set_camera_world(cam)     # (camera node world matrix, PE.cpp:860-861)
                e.update(DT)
                e.draw()
                units += e.getParticlesCount()   # device -> host: live count (= quads drawn) of this frame
                continue
            # what an engine's particle system calls once per frame: transforms in, (spawns,) update + draw, live counts out
            if churn:
                ctx.batch_set_transforms(emitters, nodes, cam, handles)
                ctx.batch_emit_once(emitters, [spawn_per_frame] * len(emitters), handles)
            nw, cw = (None, None) if churn else (nodes, cam)
            if pipelined:
                # the counts of frame s - 1 arrive with call s: the host never waits for the frame it has just enqueued
                if ctx.batch_frame_pipelined(emitters, DT, nw, cw, handles, e2e_counts):
                    units += int(e2e_counts.sum())
            else:
                ctx.batch_frame(emitters, DT, nw, cw, handles, e2e_counts)
                units += int(e2e_counts.sum())
        if pipelined and not single:
            units += int(ctx.batch_counts(emitters, handles).sum())   # the last frame's counts (waits for it)
        ctx.sync()
        dt = time.perf_counter() - t0
        b1 = ctx.transfer_bytes()
        return units, dt, (b1[0] - b0[0]) / e2e_steps, (b1[1] - b0[1]) / e2e_steps

    e2e_blocking = None
    if not single:
        u, t, _, _ = e2e_run(False)
        e2e_blocking = (u, t)
    e2e_units, t_e2e, h2d_step, d2h_step = e2e_run(True)
    matrix_bytes = 2 * 64 * len(emitters)   # the two float[16] the caller hands over per emitter per frame

    # ---- e2e with the vertex stream delivered to pinned HOST memory every frame: what a renderer that submits from client
    #      arrays (the reference's glDrawElements path, MeshBatch.cpp:299-303) would need.  Frame k is drawn straight into
    #      device buffer k mod 2 (t3d_emitter_draw_to) and travels over PCIe on the download stream while frame k + 1 is
    #      computed; the figure is PCIe-bound (144 B per quad). ----
    e2e_v = None
    if not args.no_vertices_to_host and workload == "c3" and world == 1:
        e2e_v = vertices_to_host(ctx, emitters[0], local_particles, torch)

    # ---- reduce over ranks: max time, sum of work (the only collective, after the run) ----
    from tractor3d_b200 import sharding
    mx, sm = sharding.reduce_stats({"ms_total": ms_total, "t_e2e": t_e2e, "t_e2e_blocking": e2e_blocking[1] if e2e_blocking else 0.0},
                                   {"updates": updates_done, "e2e_units": e2e_units, "particles": local_particles,
                                    "live": live_after, "launches": launches, "e2e_blocking_units": e2e_blocking[0] if e2e_blocking else 0})
    ms_total, t_e2e = mx["ms_total"], mx["t_e2e"]
    updates_done, e2e_units, total_particles, live_after, launches = sm["updates"], sm["e2e_units"], sm["particles"], sm["live"], sm["launches"]

    if rank == 0:
        per_launch_particles = local_particles if not churn else local_updates_done / args.steps
        moved = spawn_per_frame * len(emitters) if churn else 0
        roof = kernel_rooflines(kern, per_launch_particles, animated_frac, moved, moved, measured_traffic().get(workload))
        dominant = max(roof, key=lambda k: kern[k]["ms_total"]) if roof else None
        value = updates_done / (ms_total * 1e-3)
        line = {
            "metric": "particle updates/s (one step = update + billboard draw of every live particle; quads/s is the same number)",
            "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            # c4 / c2: a fixed population split by emitter ownership (strong); c3 / c3d: one emitter per GPU, replicas only;
            # c5: 1 Mi spawned per frame per GPU (weak).  The default N > 1 workload is c4: its N = 1 point is `also.c4`
            # of the N = 1 line (the N = 1 headline is c3, the configuration the metric is quoted on).
            # (the default command's N > 1 runs measure c4, strong scaling; so that the series reads consistently the N = 1
            #  line says "strong" too and names where its own point of that series is: also.c4)
            "scaling": "strong" if (workload in ("c4", "c2") or (workload == "c3" and args.workload is None)) else "weak",
            "scaling_note": ("N = 1 headline is c3 (one emitter cannot be sharded); the N > 1 runs of this command measure c4, whose N = 1 "
                             "point is also.c4.value of this line; also.c5 is weak-scaled (1 Mi spawned per frame per GPU)")
                            if (workload == "c3" and args.workload is None) else None,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl_name, "particles_total": total_particles, "live_after": live_after,
                       "dt_ms": DT, "rng": "device counter-based generator (T3D_RNG_DEVICE)",
                       "rotation_mode": "frozen (reference-identical, SB.cpp:339)" if args.rotation == "frozen"
                                        else "per particle (createRotation(right x up, angle_i) for every quad; not what the reference does)",
                       "l2": "inputs larger than L2 (state + vertex stream >> 126 MB), no flush needed" if total_particles // world * 280 > 512 * MI
                             else "working set may fit L2; no flush",
                       "parallelism": f"emitter-ownership sharding x{world}, no per-frame collective"},
            "quads_per_s": value,
            "roofline": roof.get(dominant), "roofline_dominant_kernel": dominant, "rooflines": roof,
            "kernels": kern, "gpu_launches": launches,
            "e2e": {"value": e2e_units / t_e2e, "unit": "particles/s",
                    "h2d_bytes_per_step": h2d_step + matrix_bytes, "d2h_bytes_per_step": d2h_step,
                    "steps": e2e_steps,
                    "blocking_counts": (sm["e2e_blocking_units"] / mx["t_e2e_blocking"]) if mx["t_e2e_blocking"] > 0 else None,
                    "what": "one emitter: set_node_world/set_camera_world/update/draw/getParticlesCount calls, the count read back before the "
                            "next frame is issued; many emitters: t3d_batch_frame_pipelined per frame (node and camera matrices in, update + "
                            "draw, the live counts of every frame read back -- call k hands over those of frame k - 1, so the host prepares "
                            "frame k + 1 while frame k runs; blocking_counts = the same with t3d_batch_frame, which waits for its own "
                            "frame's counts); "
                            "the vertex stream stays in HBM for a renderer that takes the device pointer or a mapped interop "
                            "buffer (t3d_emitter_vertices / t3d_emitter_draw_to); see e2e_vertices_to_host for the host-buffer figure"},
            "e2e_vertices_to_host": e2e_v,
            "clocks": clocks,
        }
        also_plan = []
        if not args.no_also and workload == "c3" and world == 1:
            also_plan = [("c2", 50), ("c3d", 30), ("c5", 40), ("c4", 12)]     # every BASELINE config on one GPU
    else:
        also_plan = []
    if world > 1 and not args.no_also and workload == "c4":
        also_plan = [("c5", 40)]                                              # 1 Mi spawned per frame PER GPU
    also = {}
    if also_plan:
        for e in emitters:
            e.release()
        emitters = []
        for wl, st in also_plan:
            barrier()
            r = quick_measure(ctx, stream, wl, st, torch, rank, world)
            if world > 1:
                mx, sm = sharding.reduce_stats({"ms": r["ms_per_step"]}, {"particles": r["particles"]})
                r["ms_per_step_max_over_ranks"] = mx["ms"]
                r["particles_all_ranks"] = sm["particles"]
                r["value"] = sm["particles"] / (mx["ms"] * 1e-3)
                r["note"] = "value = live particles of all ranks / slowest rank's step time; kernels and rooflines are rank 0's"
            also[wl] = r
    if rank == 0:
        if also:
            line["also"] = also
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single_core(args.cpu_particles)
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


# --------------------------------------------------------------------------------------------------
# CPU legs: the reference's own update()/draw() (oracle/_ref) or, where that was never built, the oracle port
# --------------------------------------------------------------------------------------------------
def _cpu_lib():
    import cpu_harness as H
    lib = H.ref(fresh=True)
    return lib if lib is not None else H.oracle(fresh=True)


def _cpu_worker(n_particles, frames, seed, q=None):
    from tractor3d_b200 import presets
    lib = _cpu_lib()
    lib.rand_seed(seed)
    p, sprite = presets.headline_64m(capacity=n_particles)
    e = lib.emitter(p)
    e.set_sprite_frame_coords(*sprite)
    e.emit_once(n_particles)
    e.update(DT); e.draw()                      # warm-up frame (also latches the frozen rotation)
    t_u = lib.time_update(e.h, DT, frames)
    t_d = lib.time_draw(e.h, frames)
    out = (n_particles * frames, t_u, t_d, lib.kind)
    if q is not None:
        q.put(out)
    return out


def cpu_baseline_single_core(n_particles):
    frames = 30
    units, t_u, t_d, kind = _cpu_worker(n_particles, frames, 1)
    return {"value": units / (t_u + t_d), "unit": "particles/s", "cores": 1, "kind": kind,
            "sample": f"c3 parameters, 1 emitter x {n_particles} particles x {frames} frames of update()+draw(), one host core "
                      f"(the reference path is single-threaded and non-reentrant)",
            "update_only": units / t_u, "draw_only": units / t_d}


def _cpu_worker_steps(n_particles, frames, seed, n_steps, q):
    """One forked process: builds its emitter once, then times n_steps steps of `frames` update()+draw() frames."""
    from tractor3d_b200 import presets
    lib = _cpu_lib()
    lib.rand_seed(seed)
    p, sprite = presets.headline_64m(capacity=n_particles)
    e = lib.emitter(p)
    e.set_sprite_frame_coords(*sprite)
    e.emit_once(n_particles)
    e.update(DT); e.draw()                      # also latches the frozen rotation
    times = []
    for _ in range(n_steps):
        times.append(lib.time_frames(e.h, DT, frames))
    q.put((times, lib.kind))


def run_reference(args):
    """The reference's own CPU path (oracle/_ref: the unmodified ParticleEmitter/SpriteBatch code compiled headless; the
    plain-C port only where that library was never built) on every host core.  The path is single-threaded and
    non-reentrant (function-local statics, rand()), so parallelism is one forked process per core, each with its own
    emitter -- the same emitter-ownership sharding the GPUs use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = min(len(os.sched_getaffinity(0)), 256)
    n_particles = max(32 * 1024, args.cpu_particles // 4)
    frames = 4
    n_steps = args.warmup + args.steps
    ctxm = mp.get_context("fork")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_cpu_worker_steps, args=(n_particles, frames, 100 + k, n_steps, q)) for k in range(cores)]
    for p in procs:
        p.start()
    outs = [q.get() for _ in procs]
    for p in procs:
        p.join()
    kind = outs[0][1]
    # per step: every process advances n_particles x frames; the step lasts as long as the slowest process
    step_s = [max(o[0][k] for o in outs) for k in range(args.warmup, n_steps)]
    units = cores * n_particles * frames
    results = [units / t for t in step_s]
    value = statistics.median(results)
    line = {
        "impl": "reference",
        "metric": "particle updates/s (one step = update + billboard draw of every live particle; quads/s is the same number)",
        "value": value, "unit": "particles/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": statistics.median(step_s) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"c3 parameters; bounded sample: {cores} forked processes (one per host core; threads are invalid: "
                               f"function-local statics and rand()) x 1 emitter x {n_particles} particles x {frames} frames per step"},
        "cpu_baseline": {"value": value, "unit": "particles/s", "cores": cores, "kind": kind,
                         "sample": f"{cores} processes x {n_particles} particles x {frames} frames of update()+draw() per step"},
        "e2e": {"value": value, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)


_RESULT_OUT = None


def emit_line(line):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries write banners to file descriptor 1 (NCCL prints its version there at communicator creation): keep
    # stdout for the result line alone and send everything else to stderr.
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default=None, choices=["c2", "c3", "c3d", "c4", "c5"])
    ap.add_argument("--particles", type=int, default=0, help="override the population size of the workload")
    ap.add_argument("--cpu-particles", type=int, default=2 * MI, help="population of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the extra c2 measurement")
    ap.add_argument("--rotation", default="frozen", choices=["frozen", "per-particle"],
                    help="billboard rotation: the reference's frozen matrix (default) or a matrix per quad (experiments)")
    ap.add_argument("--churn-spawn", type=int, default=None, help="c5: particles spawned per frame per GPU")
    ap.add_argument("--churn-energy", default=None, help="c5: energy range in ms, e.g. 500,2000")
    ap.add_argument("--no-vertices-to-host", action="store_true",
                    help="skip the frames that also stream the vertices to pinned host memory (PCIe-bound e2e figure)")
    args = ap.parse_args()
    if args.churn_spawn:
        CHURN["c5"]["spawn"] = args.churn_spawn
    if args.churn_energy:
        CHURN["c5"]["energy"] = tuple(float(x) for x in args.churn_energy.split(","))
        CHURN["c5"]["warm"] = max(30, int(CHURN["c5"]["energy"][1] / DT) + 10)
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args)
    else:
        args.steps = args.steps if args.steps is not None else 100
        args.warmup = args.warmup if args.warmup is not None else 10
        run_cuda(args)


if __name__ == "__main__":
    main()
