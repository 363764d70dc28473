#!/usr/bin/env python
"""bench.py — decoded MVerts/s of batched .crt decode on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--batch 256] [--impl ours|reference]

A "step" is one pass of the decode hot path over one batch of synthetic meshes (BASELINE configs[1]: 256 x 128K-vertex
pos14/uv12/normal10 meshes per GPU).  Work shards by mesh: every rank decodes its own batch (weak scaling), there is no
data-path collective inside the decode; torch.distributed carries the barrier, the max-over-ranks time and — in
the `shard` block — the one exchange the path has: the scatter of the compressed blobs from the ingest rank.

  value     verts decoded by all ranks / max-over-ranks device time, blobs already resident in HBM, outputs left in HBM.
            The host directory walk (O(#blocks), microseconds per mesh) + its H2D is re-done INSIDE every timed step.
  e2e       same metric through the public API with HOST buffers: H2D of the blobs from pinned memory, kernels,
            D2H of every output arena into pinned memory, all inside the timed region.
  roofline  dominant kernel (by measured device time): algorithmic bytes (blob + bound outputs) / its mean
            duration (the library's CUDA-event stage timers over a re-run of the same steps right after the timed region)
            against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the unmodified reference (oracle/_ref) timed on this box's host cores, rank 0, N=1 only.
  shard     BASELINE configs[3] as north_star words it (every N): 4096 mixed meshes held by rank 0 in its HBM -> LPT plan from the
            walk tapes -> ONE grouped NCCL send/recv into each rank's device arena -> every rank decodes its bin; scatter_ms,
            decode_ms, LPT max/mean load, strong-scaling value; the gather of one output arena timed separately.
  secondary (N=1) tarta x 64, configs[4], configs[2] device-resident, and the latency of ONE configs[0] mesh through the
            crt::Decoder-shaped host call next to the reference's.

Inputs: the synthetic meshes are ENCODED by the reference's own Encoder (oracle/workloads.py -> oracle/_ref; this repo has no
encoder, SURVEY §8 puts it out of scope) as set-up before anything is timed; without oracle/_ref one pre-encoded blob per
workload from tests/golden/bench/ is replicated and `data` says so.  Nothing under oracle/ runs inside a timed region of the
default arm except the cpu_baseline leg; the decode itself goes through corto_b200 (C ABI) only.

`--impl reference` times the reference's own single-threaded C++ decoder, one Decoder per host thread over all host
threads, on the same workload (bounded sample per step).
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

import numpy as np  # noqa: E402


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "tarta"])
    p.add_argument("--batch", type=int, default=None, help="meshes per GPU (default: 256 for c2, 512 c3, 512 c4, 1 c1/c5)")
    p.add_argument("--distinct", type=int, default=None, help="distinct seeds to encode (default: all)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-shard", action="store_true", help="skip the configs[3] scatter + sharded decode block")
    p.add_argument("--no-secondary", action="store_true", help="skip the tarta / c5 / c3 secondary lines (N=1 only)")
    p.add_argument("--shard-batch", type=int, default=4096, help="meshes of the configs[3] batch the ingest rank holds")
    p.add_argument("--shard-distinct", type=int, default=256, help="distinct seeds encoded for it (the rest are repeats)")
    return p.parse_args()


def measured_traffic(kernel, workload, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full capture of the SAME launch
    configuration (profiles/r2_traffic_c2.json: 256 x configs[1]); None for any other kernel / workload / batch — never a constant."""
    if workload != "c2" or batch != 256:
        return None
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic_c2.json")))["kernels"]
        for name, v in d.items():
            if name.split("<")[0] == kernel.split("<")[0].split(" ")[0]:
                return int(v["dram_bytes_read"] + v["dram_bytes_write"])
    except Exception:
        pass
    return None


DEFAULT_BATCH = dict(c1=1, c2=256, c3=512, c4=512, c5=1, tarta=64)


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def algorithmic_bytes(bd):
    """len(blob) + every bound output array."""
    out = sum(int(t.numel()) * t.element_size() for t in bd.out.values())
    return int(bd.total_bytes), out


def run_reference(args, rank, world, blobs):
    """Reference arm: the unmodified reference's CPU decode, all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import refshim, pyoracle
    threads = os.cpu_count() or 1
    sample = blobs[:max(1, min(len(blobs), 4 * threads))]
    verts = sum(pyoracle.info(b)["nvert"] for b in sample)
    kind = "reference" if refshim.available() else "port"

    def one():
        if kind == "reference":
            return refshim.decode_bench(sample, threads, 1)
        t0 = time.perf_counter()
        pyoracle.decode_all(sample)
        return time.perf_counter() - t0
    for _ in range(args.warmup):
        one()
    ts = [one() for _ in range(args.steps)]
    total = sum(ts)
    val = verts * args.steps / total / 1e6
    desc = "%d of %d meshes per step" % (len(sample), len(blobs))
    print(json.dumps({
        "impl": "reference", "metric": "decoded MVerts/s (batched .crt)", "value": val, "unit": "MVerts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32+fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME(args), "batch_per_gpu": len(blobs), "sample": desc},
        "cpu_baseline": {"value": val, "unit": "MVerts/s", "cores": threads if kind == "reference" else 1, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": "MVerts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_shard(args, rank, world, local):
    """BASELINE configs[3] as north_star words it: `--shard-batch` mixed meshes held by rank 0 (in its HBM) -> LPT plan from the
    walk tapes -> ONE grouped NCCL send/recv into each rank's device arena -> every rank decodes its bin (directory rebuilt from
    the tapes, no payload byte returns to a host).  Strong scaling: the batch is fixed, N grows.  Device time, max over ranks."""
    import torch
    import torch.distributed as dist
    import corto_b200
    from corto_b200 import dist as cd
    from oracle import workloads
    dev = torch.device("cuda", local)
    ing = None
    if rank == 0:
        blobs = workloads.build("c4", args.shard_batch, seed0=1, distinct=min(args.shard_distinct, args.shard_batch))
        ing = cd.Ingest(blobs, world, device=dev, pin=True)           # set-up: the batch sits in rank 0's HBM, bin after bin
        del blobs
    if world > 1:
        got = cd.scatter_blobs(ing, src=0, device=dev)
    else:
        m = ing.meta()
        got = dict(arena=ing.arena, tapes=[np.frombuffer(t, dtype=np.uint8) for t in m["tapes"]], lens=m["lens"], ids=m["ids"][0], meta=m)
        got["tapes"] = [got["tapes"][i] for i in got["ids"]]; got["lens"] = [m["lens"][i] for i in got["ids"]]
    bd = corto_b200.BatchDecoder.from_device(got["tapes"], got["lens"], got["arena"])
    bd.allocate()
    bd.upload()
    bd.decode()
    torch.cuda.synchronize()
    rc, _ = bd.status()
    assert rc == 0, "sharded decode failed: %s" % corto_b200.lib().crt_last_error()
    steps, sc, de, tot = max(2, min(args.steps, 5)), [], [], []
    for k in range(steps + 1):                                            # first pass = warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        if world > 1:
            cd.scatter_blobs(ing, src=0, device=dev, meta=got["meta"], out=got["arena"])
        e1.record()
        bd.rewalk(); bd.decode()
        e2.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2), e0.elapsed_time(e2)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if k:
            sc.append(float(t[0])); de.append(float(t[1])); tot.append(float(t[2]))
    rc, _ = bd.status()
    assert rc == 0
    # optional gather of one output arena (positions) to rank 0: bounded by rank 0's NVLink ingress, timed separately
    gather_ms = None
    if world > 1:
        rows = [None] * world
        dist.all_gather_object(rows, int(bd.total_verts))
        g = cd.gather_rows(bd.out["position"], rows, dst=0)          # first pass: NCCL opens the reverse channels, the allocator grows
        del g
        dist.barrier(); torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        g = cd.gather_rows(bd.out["position"], rows, dst=0)
        g1.record(); torch.cuda.synchronize()
        t = torch.tensor([g0.elapsed_time(g1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather_ms = float(t[0])
        del g
    verts = torch.tensor([float(bd.total_verts)], device="cuda")
    if world > 1:
        dist.all_reduce(verts)
    meta = got["meta"]
    sent = sum(meta["slice_len"][1:]) if world > 1 else 0
    ms = float(np.mean(tot))
    return {"workload": "c4: %d mixed meshes 8K-256K verts, all attributes + groups (%d distinct seeds), held by rank 0, LPT-sharded" % (args.shard_batch, min(args.shard_distinct, args.shard_batch)),
            "scaling": "strong", "n_gpus": world, "value": float(verts[0]) / (ms * 1e-3) / 1e6, "unit": "MVerts/s", "ms_per_step": ms,
            "scatter_ms": float(np.mean(sc)), "decode_ms": float(np.mean(de)), "steps": steps,
            "scatter_bytes": int(sent), "scatter_gbs": (sent / (np.mean(sc) * 1e-3) / 1e9) if sent else None,
            "lpt_max_over_mean_load": meta["load_max_over_mean"], "gather_position_ms": gather_ms,
            "exchange": "one torch.distributed.batch_isend_irecv (ncclGroupStart/End) from rank 0; tapes as metadata; no host bounce",
            "limiter": "rank 0's NVLink egress for scatter_ms; the largest LPT bin's decode for decode_ms"}


def run_secondary(args):
    """N=1 only: the other BASELINE configs and the real scan, device-resident, a few steps each (driver-visible side lines)."""
    import torch
    import corto_b200
    from oracle import workloads, refshim
    out = {}
    for w, batch, distinct in (("tarta", 64, 1), ("c5", 1, 1), ("c3", 512, 16)):
        if w == "tarta" and not os.path.exists(refshim.TARTA):
            continue
        try:
            blobs = workloads.build(w, batch, seed0=1, distinct=distinct)
            bd = corto_b200.BatchDecoder(blobs)
            bd.allocate(); bd.upload()
            for _ in range(2):
                bd.rewalk(); bd.decode()
            torch.cuda.synchronize()
            rc, _ = bd.status()
            assert rc == 0
            k = 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                bd.rewalk(); bd.decode()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / k
            ib, ob = algorithmic_bytes(bd)
            out[w] = {"workload": workloads.DESCRIPTION[w], "batch": batch, "value": bd.total_verts / (ms * 1e-3) / 1e6, "unit": "MVerts/s",
                      "ms_per_step": ms, "step_achieved_gbs": (ib + ob) / (ms * 1e-3) / 1e9}
            del bd, blobs
            torch.cuda.empty_cache()
        except Exception as e:                                            # a side line must never take the headline down
            out[w] = {"error": str(e)[:200]}
    # configs[0]: ONE 34K-vertex mesh through the crt::Decoder-shaped call with host buffers (latency, not throughput)
    try:
        blob = workloads.build("c1", 1)[0]
        d = corto_b200.Decoder(blob); d.decode()
        ts = []
        for _ in range(5):
            d = corto_b200.Decoder(blob)
            t0 = time.perf_counter(); d.decode(); ts.append(time.perf_counter() - t0)
        lat = {"workload": workloads.DESCRIPTION["c1"], "ours_ms": 1e3 * min(ts)}
        if refshim.available():
            lat["reference_1thread_ms"] = 1e3 * refshim.decode_bench([blob], 1, 5)
        out["c1_latency"] = lat
    except Exception as e:
        out["c1_latency"] = {"error": str(e)[:200]}
    return out


def WORKLOAD_NAME(args):
    from oracle import workloads
    return "%s: %s" % (args.workload, workloads.DESCRIPTION[args.workload])


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    batch = args.batch or DEFAULT_BATCH[args.workload]

    from oracle import workloads, refshim
    if args.impl == "reference":
        if rank == 0:
            blobs = workloads.build(args.workload, batch, seed0=1, distinct=args.distinct or min(batch, 32))
            run_reference(args, rank, world, blobs)
        return

    import torch
    import torch.distributed as dist
    import corto_b200
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # ---- synthetic batch for this rank (different seeds per rank): set-up, not timed -------------------------------
    blobs = workloads.build(args.workload, batch, seed0=1 + rank * batch, distinct=args.distinct)
    # one pinned host buffer holds every blob (16-byte aligned slices): the e2e leg's H2D source
    offs, tot = [], 0
    for b in blobs:
        offs.append(tot)
        tot += (len(b) + 15) // 16 * 16
    pinned = torch.empty(tot + 16, dtype=torch.uint8).pin_memory()
    pn = pinned.numpy()
    base = (-pn.ctypes.data) % 16
    views = []
    for b, o in zip(blobs, offs):
        v = pn[base + o: base + o + len(b)]
        v[:] = b
        views.append(v)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident decode: `value` ---------------------------------------------------------------------------
    bd = corto_b200.BatchDecoder(views)
    bd.allocate()
    bd.upload()
    for _ in range(max(args.warmup, 3)):
        bd.rewalk(); bd.decode()
    torch.cuda.synchronize()
    rc, st = bd.status()
    assert rc == 0, "decode failed: %s" % corto_b200.lib().crt_last_error()
    in_bytes, out_bytes = algorithmic_bytes(bd)
    sampler = ClockSampler(local)
    stage_acc = {}
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        bd.rewalk()          # host directory walk + H2D of the directory: part of Decoder::decode, so part of the step
        bd.decode()
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms = ev0.elapsed_time(ev1)
    tms = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    total_verts = bd.total_verts * world
    value = total_verts * args.steps / (ms_max * 1e-3) / 1e6
    # Per-kernel durations (roofline): the same steps once more with the stage timers on (the timed region above runs without
    # them and without a host sync per step, so that the directory walk of step k+1 overlaps the kernels of step k).
    bd.set_profiling(True)
    per_step_stage = []
    for _ in range(min(args.steps, 5)):
        bd.rewalk(); bd.decode()
        torch.cuda.current_stream().synchronize()
        per_step_stage.append(bd.stage_times())
    bd.set_profiling(False)
    for st_ in per_step_stage:
        for name, t in st_:
            stage_acc.setdefault(name, []).append(t)
    stage_ms = {k: float(np.mean(v)) for k, v in stage_acc.items()}
    dom = max(stage_ms, key=stage_ms.get) if stage_ms else None
    peak, peak_src = peaks()
    roof = None
    if dom:
        ach = (in_bytes + out_bytes) / (stage_ms[dom] * 1e-3) / 1e9
        kernel_of = {"clers": "k_clers_cta (+ k_clers_lf for irregular meshes)", "delta": "k_delta_mesh_seg", "cloud_fused": "k_cloud_chain / k_unpack_fused<CLOUD>",
                     "tun_decode": "k_tun_decode", "bit_unpack": "k_unpack_chain / k_unpack_fused<MESH>", "normals": "k_adj_build + k_normal_estimate",
                     "dequant": "k_dequant", "tun_tables": "k_tun_tables"}
        traffic = measured_traffic(kernel_of.get(dom, dom), args.workload, batch)
        roof = {"bound": "hbm", "kernel": kernel_of.get(dom, dom), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel_ms": stage_ms[dom], "stage_ms": stage_ms,
                "algorithmic_bytes": {"blobs": in_bytes, "outputs": out_bytes},
                "step_achieved": (in_bytes + out_bytes) * args.steps / (ms * 1e-3) / 1e9,
                "read_only_frac": in_bytes / (stage_ms[dom] * 1e-3) / 1e9 / peak,
                "stage_ms_from": "a re-run of the same steps right after the timed region with the library's stage timers (CUDA events) on",
                "traffic_source": "profiles/r2_traffic_c2.json (ncu --set full of this launch configuration)" if traffic else None,
                "note": "no kernel of the mesh path is HBM-bound except the normal estimation (61 % of peak DRAM throughput); the others are bound by "
                        "dependent-instruction latency at low occupancy (CLERS: one CTA per mesh) or by issue rate; see DESIGN.md sections 4-5"}
    launches = bd.launches * args.steps
    verts_per_gpu, faces_per_gpu = int(bd.total_verts), int(bd.total_faces)

    # ---- end to end through the public API with host buffers ---------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # Two batches in flight on two streams (double buffering): while batch k's outputs travel D2H, batch k+1 is uploaded
        # and decoded.  Every step still does H2D + decode + D2H of ITS batch inside the timed region.
        bd.set_profiling(False)
        bd2 = corto_b200.BatchDecoder(views)
        bd2.allocate()
        bds = [bd, bd2]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        host_out = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in b.out.items()} for b in bds]
        d2h = sum(int(t.numel()) * t.element_size() for t in host_out[0].values())

        def e2e_step(k):
            with torch.cuda.stream(streams[k]):
                bds[k].upload()                               # directory walk + H2D of blobs (pinned) and tables
                bds[k].decode()
                for name, v in bds[k].out.items():
                    host_out[k][name].copy_(v, non_blocking=True)   # D2H of every output arena
        for s_ in range(4):
            e2e_step(s_ & 1)
        barrier()
        ke = max(4, min(args.steps, 10))
        t0 = time.perf_counter()
        for s_ in range(ke):
            e2e_step(s_ & 1)
        for st_ in streams:
            st_.synchronize()
        barrier()
        dt = time.perf_counter() - t0                         # copies span two streams: wall clock around a full sync on both sides
        ems = torch.tensor([dt * 1e3], device="cuda")
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        rc2, _ = bd2.status()
        assert rc2 == 0
        e2e = {"value": total_verts * ke / (float(ems.item()) * 1e-3) / 1e6, "unit": "MVerts/s", "h2d_bytes_per_step": int(in_bytes),
               "d2h_bytes_per_step": int(d2h), "steps": ke, "pipelining": "2 batches in flight on 2 streams; wall clock over a full device sync"}
        # the host ceiling beside it: a plain device -> pinned-host copy of one output arena set, nothing else running
        big = max(host_out[0].values(), key=lambda t: t.numel() * t.element_size())
        src = bds[0].out[[k for k, v in host_out[0].items() if v is big][0]]
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            big.copy_(src, non_blocking=True)
        c1.record(); torch.cuda.synchronize()
        d2h_gbs = 3 * big.numel() * big.element_size() / (c0.elapsed_time(c1) * 1e-3) / 1e9
        e2e["d2h_pinned_gbs"] = d2h_gbs
        e2e["d2h_bound_mverts_s"] = world * d2h_gbs * 1e9 / (d2h / max(bds[0].total_verts, 1)) / 1e6   # if the step were nothing but its D2H

    clocks = sampler.stop()          # sampled across both timed regions (device-resident and end-to-end)

    # ---- configs[3]: scatter over NVLink + sharded decode (strong scaling), and the side lines ---------------------------
    del bd
    if not args.no_e2e:
        del bd2, bds, host_out
    torch.cuda.empty_cache()
    shard = None
    if not args.no_shard:
        try:
            shard = run_shard(args, rank, world, local)
        except Exception as e:
            shard = {"error": str(e)[:300]}
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        secondary = run_secondary(args)

    # ---- CPU baseline beside it (rank 0, N=1) -----------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import pyoracle
        sample = blobs[:min(len(blobs), 64)]
        sv = sum(pyoracle.info(b)["nvert"] for b in sample)
        if refshim.available():
            t1 = refshim.decode_bench(sample, 1, 3)
            tall = refshim.decode_bench(sample, os.cpu_count() or 1, 3)
            cpu = {"value": sv / t1 / 1e6, "unit": "MVerts/s", "cores": 1, "kind": "reference",
                   "sample": "%d of %d meshes, best of 3 passes, Decoder ctor+set*+decode()" % (len(sample), len(blobs)),
                   "all_cores": {"value": sv / tall / 1e6, "cores": os.cpu_count()}}
        else:
            t0 = time.perf_counter(); pyoracle.decode_all(sample); t1 = time.perf_counter() - t0
            cpu = {"value": sv / t1 / 1e6, "unit": "MVerts/s", "cores": 1, "kind": "port", "sample": "%d of %d meshes, 1 pass" % (len(sample), len(blobs))}

    if rank == 0:
        print(json.dumps({
            "metric": "decoded MVerts/s (batched .crt)", "value": value, "unit": "MVerts/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32+fp32", "data": "synthetic" if refshim.available() else "synthetic (one pre-encoded blob replicated: reference encoder not built here)",
            "config": {"workload": WORKLOAD_NAME(args), "batch_per_gpu": batch, "verts_per_gpu": verts_per_gpu,
                       "faces_per_gpu": faces_per_gpu, "blob_bytes_per_gpu": int(in_bytes), "output_bytes_per_gpu": int(out_bytes),
                       "parallelism": "mesh-sharded x%d, no data-path collective" % world,
                       "l2": "inputs+outputs (%.2f GB) larger than the 126 MB L2; no explicit flush" % ((in_bytes + out_bytes) / 1e9),
                       "timed_region": "host directory walk + H2D directory + all decode kernels"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "shard": shard, "secondary": secondary, "wall_s": t_wall}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
