#!/usr/bin/env python
"""bench.py -- headline benchmark of the integrating-sphere hot path (BASELINE.json config C3).

A "step" = one pass of the hot path over one batch of synthetic source rays:
trace (source -> multi-bounce loop with the CustomMirror BRDF) + 180x90 flux map + (N>1) one all-reduce.
Metric = ray-bounces/s of the whole job.  Default "scaling": "strong" -- BASELINE.json configs[2]: the SAME
1e9 rays per step on 1/2/4/8 GPUs (the reference's analogue is SetMaxThreads on a fixed ray count,
fluxAtObserverFast.C:1083-1087); --scaling weak gives every GPU --rays rays per step.  The line carries both rates
(`other_scaling`) and `map_crc`, the CRC-32 of the all-reduced map of a fixed 1e7-ray probe: equal at every N.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--rays R] [--scaling strong|weak] [--map direction|line]
                  [--impl reference]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ray_bounces_per_s_fluxmap"
UNIT = "ray-bounces/s"
FLOP_PER_BOUNCE = 100.0          # SURVEY.md 8(d)'s per-unit figure (Lambert + Gaussian roughness): the contract's unit of work
FLOP_PER_BOUNCE_MODEL = 175.0    # what the CustomMirror step executes (FMA = 2), from the per-function ncu table in profiles/ 
REF_RECORDED = 3.7e6             # reference's own recorded rate, author's PC, <=4 threads
C3_RAYS = 1_000_000_000          # BASELINE.json configs[2]
PROBE_RAYS = 10_000_000          # map_crc probe


def workload_scene(mod):
    # C3: fluxAtObserverFast.C:33-41 scene + per-bounce spec/diffuse mixture of nonLambertianFlux.C:147-211
    return mod.scene(theta_max=170.0, world_half=300.0, reflectance=0.99, roughness=0.01, max_bounces=50000,
                     brdf_kind=1, brdf_param=(0.3, 0.4, 0.6, 0.0))


def workload_config(args, mode_name, world):
    total = args.rays if args.scaling == "strong" else args.rays * world
    return {"workload": "C3 nonLambertianFlux: CustomMirror BRDF (0.3,0.4,0.6), theta_max=170, rho=0.99, "
                        "sigma=0.01, src(-60,0,-75) dir(5,0,0), 180x90 map",
            "rays_per_step": total, "rays_per_gpu_per_step": total // world, "map_mode": mode_name, "seed": 4357,
            "contract": getattr(args, "contract", "fast7"),
            "l2": "no input to cache: the RNG is counter-based and every step traces NEW ray ids, so nothing a step reads "
                  "was produced by an earlier one; per-step device traffic is the map/stats buffer only"}


def host_threads():
    """Cores this process may use -- NOT omp_get_max_threads(): torch.distributed.run exports OMP_NUM_THREADS=1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def stop(self, t0=None, t1=None):
        """Samples taken inside the timed region [t0, t1] (host clock); a region shorter than the sampling period falls back
        to the samples closest to it (the sampler runs from before the warm-up on)."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        rows = [r for r in self.rows if len(r) > 3 and r[2].isdigit()]
        if t0 is not None and rows:
            inside = [r for r in rows if t0 <= r[0] <= t1 + 0.25]
            rows = inside if inside else sorted(rows, key=lambda r: abs(r[0] - t1))[:2]
        rows = [r[1:] for r in rows]
        sm = sorted(int(r[1]) for r in rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in rows if len(r) > 2 and r[2].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_rate_probe(O, sc, src, mp, threads):
    """(ray-bounces/s, bounces per ray) of the CPU path on `threads` threads: slope between two short runs, so that the
    fixed cost of a call (buffers, thread start-up: ~0.4 s) does not bias the bounded samples sized from it."""
    O.fluxmap(sc, src, 100_000, mp, seed=4357, ray_id0=1 << 41, prec=O.F64, n_threads=threads)      # cold start: threads, page faults
    pts, n = [], 50_000
    for _ in range(8):
        t0 = time.perf_counter()
        _, st = O.fluxmap(sc, src, n, mp, seed=4357, ray_id0=1 << 40, prec=O.F64, n_threads=threads)
        pts.append((time.perf_counter() - t0, st["n_bounces"], st["n_rays"]))
        if len(pts) >= 2 and pts[-1][0] - pts[-2][0] >= 0.5:
            break
        n *= 4
    (t1, b1, _), (t2, b2, r2) = (pts[-2] if len(pts) > 1 else (0.0, 0, 0)), pts[-1]
    rate = max((b2 - b1) / max(t2 - t1, 1e-3), b2 / t2)
    return rate, b2 / max(r2, 1)


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  ROOT + ROBAST cannot be built here
    (DESIGN.md), so this times the oracle's double-precision restatement on all host cores.  The thread count is
    explicit (torchrun exports OMP_NUM_THREADS=1) and every step is a bounded sample of the workload, sized from a
    half-second calibration so that warm-up + K steps take about --ref-budget seconds whatever K and the core count are."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    O.build()
    mode = O.MAP_DIRECTION if args.map == "direction" else O.MAP_LINE
    sc, src, mp = workload_scene(O), O.source(), O.map_spec(mode=mode)
    cores = host_threads()
    if args.ref_rays > 0:
        sample = args.ref_rays
    else:
        rate, bpr = cpu_rate_probe(O, sc, src, mp, cores)
        per_step = args.ref_budget / (args.steps + 0.1 * args.warmup)
        sample = int(min(max(rate * per_step / bpr, 2_000), 50_000_000))
    for w in range(args.warmup):
        O.fluxmap(sc, src, max(sample // 10, 1000), mp, seed=4357, ray_id0=w * sample, prec=O.F64, n_threads=cores)
    bounces = 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        _, st = O.fluxmap(sc, src, sample, mp, seed=4357, ray_id0=(args.warmup + s) * sample, prec=O.F64, n_threads=cores)
        bounces += st["n_bounces"]
    dt = time.perf_counter() - t0
    v = bounces / dt
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = workload_config(args, args.map, world)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} rays per step of the same workload ({bounces} bounces in {dt:.1f} s on "
                                       f"{cores} threads); ROOT+ROBAST are not installable here, this is the FP64 oracle restatement"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "rays_per_s": sample * args.steps / dt, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    # stdout carries exactly ONE line, the JSON: libraries that write to fd 1 on their own (NCCL prints its version banner
    # at the first communicator when NCCL_DEBUG is set on the box) are sent to stderr until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def profile_reference(contract):
    """ncu-derived numbers are NOT measured by this run: they are read from the committed capture summary
    (profiles/rNN_k_trace_ncu_<contract>.json, written by tools/ncu_summary.py --json) together with the capture's own
    configuration."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_k_trace_ncu_{contract}.json")))
    if not files:
        return None
    try:
        with open(files[-1]) as f:
            d = json.load(f)
        d["file"] = os.path.relpath(files[-1], ROOT)
        return d
    except Exception:
        return None


def _run_ours(args):
    import ctypes as C
    import zlib
    import numpy as np
    import torch
    import torch.distributed as dist
    import altair_raytracing_b200 as A
    from altair_raytracing_b200.distributed import ShardedTracer

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = A.MAP_DIRECTION if args.map == "direction" else A.MAP_LINE
    sc, src, mp = workload_scene(A), A.source(), A.map_spec(mode=mode)
    ctx = A.Context([local])
    ctx.set_contract({"fast": A.CONTRACT_FAST, "fast7": A.CONTRACT_FAST7, "exact": A.CONTRACT_EXACT}[args.contract])
    tr = ShardedTracer(ctx, sc, src, mp, seed=4357, device=local)
    nb = mp.n_theta * mp.n_phi
    next_id = [0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(total):
        """one step of the job: `total` NEW global ray ids, this rank its shard, one all-reduce"""
        buf = tr.step_device(total, ray_id0=next_id[0])
        next_id[0] += total
        return buf

    def timed(total, steps):
        """(seconds [max over ranks], rays, bounces, launches) of `steps` device-resident steps"""
        acc = torch.zeros(8, dtype=torch.int64, device=tr.device)
        barrier()
        l0 = ctx.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            acc += step_dev(total)[nb:nb + 8]
        ev1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=tr.device)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        st = acc.cpu().numpy()
        return ms.item() * 1e-3, int(st[0]), int(st[5]), ctx.launches - l0 + (2 * steps if world > 1 else steps)

    total = args.rays if args.scaling == "strong" else args.rays * world
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step_dev(total)
    barrier()
    t_region0 = time.time()
    dt, rays_done, bounces, launches = timed(total, args.steps)
    clk = clocks.stop(t_region0, time.time()) if rank == 0 else None
    value = bounces / dt

    # ---- e2e: the public blocking call, host results every step (params H2D, map+stats D2H)
    barrier()
    t0 = time.perf_counter()
    e2e_bounces = 0
    for _ in range(args.steps):
        counts, stats = tr.step(total, ray_id0=next_id[0])
        next_id[0] += total
        e2e_bounces += int(stats[0, 5])
    barrier()
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=tr.device)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_val = e2e_bounces / e2e_dt.item()
    h2d = C.sizeof(A.Scene) + C.sizeof(A.Source) + C.sizeof(A.MapSpec) + 48
    d2h = (nb + 8) * 8

    # ---- the other scaling mode, a few steps (N = 1: the same job)
    other = None
    if world > 1 and not args.no_other:
        o_total = args.rays * world if args.scaling == "strong" else args.rays
        o_steps = min(args.steps, 3)
        step_dev(o_total)
        o_dt, o_rays, o_b, _ = timed(o_total, o_steps)
        other = {"scaling": "weak" if args.scaling == "strong" else "strong", "value": o_b / o_dt, "unit": UNIT,
                 "rays_per_step": o_total, "steps": o_steps, "ms_per_step": o_dt / o_steps * 1e3, "rays_per_s": o_rays / o_dt}

    # ---- map_crc: the all-reduced map + stats of a FIXED probe (ray ids 0 .. 1e7-1): the same bytes at every N
    c_probe, s_probe = tr.step(PROBE_RAYS, ray_id0=0)
    map_crc = "%08x" % (zlib.crc32(np.ascontiguousarray(c_probe).tobytes() + np.ascontiguousarray(s_probe[:, :6]).tobytes()) & 0xffffffff)

    # ---- N > 1: the same probe through ONE process driving all N GPUs -- altb_create(devices, N) owns NCCL communicators and
    #      merges the per-device maps with one ncclAllReduce inside the C ABI (what the C++ macros use).  Rank 0 does it while
    #      the other ranks wait at the barrier; its CRC must equal map_crc.
    inproc = None
    if world > 1 and not args.no_inproc:
        barrier()
        if rank == 0:
            try:
                t0 = time.perf_counter()
                with A.Context(list(range(world))) as many:
                    many.set_contract({"fast": A.CONTRACT_FAST, "fast7": A.CONTRACT_FAST7, "exact": A.CONTRACT_EXACT}[args.contract])
                    t1 = time.perf_counter()
                    c_in, s_in = many.trace_fluxmap(sc, src, PROBE_RAYS, mp, seed=4357, ray_id0=0)
                    t2 = time.perf_counter()
                    s_arr = np.array([[s_in[0][k] for k in ("n_rays", "n_exited", "n_exit_port", "n_absorbed", "n_suspended", "n_bounces")]], dtype=np.uint64)
                    crc_in = "%08x" % (zlib.crc32(np.ascontiguousarray(c_in).tobytes() + s_arr.tobytes()) & 0xffffffff)
                    inproc = {"devices": world, "collective": many.collective, "map_crc": crc_in, "equals_map_crc": crc_in == map_crc,
                              "create_s": t1 - t0, "probe_s": t2 - t1}
            except Exception as e:
                inproc = {"devices": world, "error": str(e)}
        barrier()

    # ---- per-kernel timing for the roofline (CUDA events inside the library, on its own stream, around the trace launches)
    l0, tl0 = ctx.launches, ctx.trace_launches
    _, kst = ctx.trace_fluxmap(sc, src, total // world, mp, seed=4357, ray_id0=next_id[0] + rank * (total // world))
    k_launches, n_trace = ctx.launches - l0, max(1, ctx.trace_launches - tl0)
    kst = kst[0]
    peak = ctx.measure_fp32_peak()
    achieved = FLOP_PER_BOUNCE * kst["n_bounces"] / kst["t_trace_s"] * 1e-12
    roofline = {"bound": "fp32", "kernel": "k_trace<rough,CustomMirror>", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": None,
                "peak_source": "FFMA-chain probe measured live on this GPU (MEASURED_PEAKS.json has no FP32 number); "
                               "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                "flop_per_bounce": FLOP_PER_BOUNCE,
                "flop_per_bounce_note": f"SURVEY 8(d) unit of work; the CustomMirror step itself executes ~{FLOP_PER_BOUNCE_MODEL:.0f} "
                                        "FP32 flop per bounce (DESIGN.md section 7), not used for `achieved`",
                "launches": n_trace, "kernel_launches_in_call": k_launches,
                "avg_launch_ms": kst["t_trace_s"] * 1e3 / n_trace, "map_ms_per_launch": kst["t_map_s"] * 1e3 / n_trace,
                "bounces_per_s_kernel": kst["n_bounces"] / kst["t_trace_s"],
                "hbm_bytes_per_bounce_algorithmic": 0.0}
    prof = profile_reference(args.contract)
    if prof and args.map == "direction":
        # the capture's DRAM bytes, scaled from the capture's launch size to this run's launch size (per launch, like `achieved`)
        if prof.get("dram_bytes_per_launch") and prof.get("bounces_in_launch"):
            roofline["traffic"] = prof["dram_bytes_per_launch"] * (kst["n_bounces"] / n_trace) / prof["bounces_in_launch"]
            roofline["traffic_note"] = ("dram__bytes_read+write of the committed ncu capture, scaled by bounces per launch; algorithmic "
                                        "HBM bytes of the in-kernel direction sink: the 129.7 kB map + statistics per launch")
        roofline["profile_reference"] = prof
    elif prof:
        roofline["profile_reference_note"] = "the committed k_trace capture is of the direction-map instance; not attached to a LINE-map run"
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.map, world),
            "rays_per_s": rays_done / dt, "bounces_per_ray": bounces / max(rays_done, 1),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "map_crc": map_crc,
            "map_crc_note": f"CRC-32 of the all-reduced map+stats of ray ids 0..{PROBE_RAYS - 1}; identical at every N",
            "other_scaling": other, "inproc_context": inproc, "roofline": roofline, "clocks": clk,
            "reference_recorded": {"value": REF_RECORDED, "unit": UNIT, "note": "BASELINE.md, author's PC, <=4 threads"}}
    if world == 1 and not args.no_cpu:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pyoracle as O
            O.build()
            osc, osrc = workload_scene(O), O.source()
            omp_ = O.map_spec(mode=O.MAP_DIRECTION if args.map == "direction" else O.MAP_LINE)
            cores = host_threads()
            rate, bpr = cpu_rate_probe(O, osc, osrc, omp_, cores)
            sample = args.ref_rays if args.ref_rays > 0 else int(min(max(rate * 10.0 / bpr, 2_000), 50_000_000))   # ~10 s
            t0 = time.perf_counter()
            _, ost = O.fluxmap(osc, osrc, sample, omp_, seed=4357, prec=O.F64, n_threads=cores)
            cdt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ost["n_bounces"] / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} rays of the same workload, FP64 oracle restatement (ROOT+ROBAST cannot "
                                              f"run here), {cdt:.1f} s on {cores} threads"}
            s1 = max(sample // (2 * max(cores, 1)), 1000)          # the same on ONE core, ~half the time again
            t0 = time.perf_counter()
            _, ost1 = O.fluxmap(osc, osrc, s1, omp_, seed=4357, prec=O.F64, n_threads=1)
            line["cpu_baseline"]["value_1core"] = ost1["n_bounces"] / (time.perf_counter() - t0)
        except Exception as e:          # the CPU leg must never cost the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rays", type=int, default=C3_RAYS,
                    help="strong: rays per step of the whole job (C3: 1e9); weak: rays per GPU per step")
    ap.add_argument("--scaling", choices=["strong", "weak"], default="strong")
    ap.add_argument("--map", choices=["direction", "line"], default="direction")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--contract", choices=["fast7", "fast", "exact"], default="fast7",
                    help="arithmetic contract of the bounce loop (include/altair_b200.h): fast = MUFU special functions, validated by "
                         "the north star's replay (<= 1e-4) and chi^2 criteria; fast7 = fast + Philox4x32-7 as the generator (another "
                         "random stream, validated statistically); exact = bit-identical to the CPU oracle")
    ap.add_argument("--ref-rays", type=int, default=0, help="CPU sample size (rays per step); 0 = sized from a calibration run")
    ap.add_argument("--ref-budget", type=float, default=90.0, help="seconds the reference arm may spend on its K steps")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the secondary scaling measurement")
    ap.add_argument("--no-inproc", action="store_true", help="N > 1: skip rank 0's in-process multi-device probe (NCCL inside the C ABI)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
