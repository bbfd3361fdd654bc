#!/usr/bin/env python
"""bench.py — converged closed-chain projections/s (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--count C] [--config NAME]

A "step" is one pass of the hot path — KinematicChainConstraint::project (ConstraintFunction.h:57-82)
batched over C synthetic Seeds-U (uniform in the joint bounds, counter-based stream) — on
every rank.  Workload at N=1: BASELINE configs[1], dumbbell, 1M seeds (fits one GPU).  N>1: weak scaling,
every rank projects its own C seeds from a disjoint counter range, then the ranks all-gather the converged
counts and the compacted converged states over NCCL (the path's one exchange step).

Printed JSON (rank 0, one line): see the keys built in main().  Definitions:
  value        converged (project()==true) projections per second, whole job, seeds resident in HBM
  e2e          same metric through the host-buffer C-ABI call (pinned host seeds in, results out)
  roofline     FP64 pipe: algorithmic FLOPs (csrc/ccp_flops.h x the kernel's own per-sample iteration
               counters) / CUDA-event time of the projection kernel, against the DFMA peak MEASURED in
               this run by ccp_fp64_peak_probe (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline the reference-faithful CPU restatement (oracle A: FD Jacobian + SVD solve, = the
               reference's arithmetic without RBDL/OMPL overhead) on all host cores, bounded sample
--impl reference times that CPU restatement alone (the reference itself cannot be built here: OMPL,
RBDL, Eigen, ROS are absent — DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "converged closed-chain projections/sec"
UNIT = "converged projections/s"


def _env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, windows=None):
        """windows: [(t0, t1)] perf_counter intervals of continuous GPU load; only samples inside one are kept."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        return self.summarise(self.lines, windows)

    @staticmethod
    def summarise(lines, windows=None):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in lines:
            if windows is not None and not any(a + 0.05 <= ts <= b for a, b in windows):
                continue
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                pw.append(float(p[3]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def _parse_cpulist(txt):
    cpus = set()
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(torch, local):
    """Pin this rank's threads (and so, by first touch, its page-locked buffers) to the NUMA node its GPU hangs off.
    Returns what was found for the JSON line; a single-node box (or a VM that hides the topology) is left alone."""
    info = {"cpus_visible": os.cpu_count(), "numa_nodes": None, "gpu_numa_node": None, "bound": False}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes"] = len(nodes)
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        info["gpu_numa_node"] = node
        if len(nodes) > 1 and node >= 0:
            cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound"] = True
        info["affinity"] = len(os.sched_getaffinity(0))
    except (OSError, ValueError, AttributeError) as e:
        info["note"] = repr(e)
    return info


def cpu_baseline(config: str, sample: int, threads: int, with_engine_arithmetic: bool = False):
    """Oracle A (reference-faithful restatement) on the first `sample` Seeds-U of the workload."""
    import numpy as np

    from closed_chain_motion_planner_b200 import grasping_point
    from oracle.oracle import OracleA

    cfg = grasping_point().loadConfig(config)
    A = OracleA(cfg.arm_indices)
    A.set_initial_position(cfg.start)
    seeds = A.seeds_uniform(0, 0, sample)
    t0 = time.perf_counter()
    r = A.project(seeds, fd=True, nthreads=threads)
    dt = time.perf_counter() - t0
    out = {"seconds": dt, "projections_per_s": sample / dt, "converged_per_s": float(r["ok"].sum()) / dt,
           "ok_fraction": float(np.mean(r["ok"])), "mean_iters": float(np.mean(r["iters"]))}
    if with_engine_arithmetic:
        # the reference runs this path on ONE thread: the same restatement on one core, smaller sample
        n1 = max(200, sample // 50)
        t0 = time.perf_counter()
        A.project(seeds[:n1], fd=True, nthreads=1)
        out["single_thread_projections_per_s"] = n1 / (time.perf_counter() - t0)
        # the engine's own arithmetic (analytic Jacobian, csrc/ccp_core.h compiled by g++) on the same cores and
        # seeds: separates the algorithmic part of the GPU/CPU ratio from the hardware part
        from closed_chain_motion_planner_b200._capi import default_model_desc
        from oracle.oracle import OracleB

        B = OracleB(default_model_desc(cfg.arm_indices))
        B.set_initial_position(cfg.start)
        t0 = time.perf_counter()
        rb = B.project(seeds, nthreads=threads)
        dtb = time.perf_counter() - t0
        out["engine_arithmetic_on_cpu"] = {"projections_per_s": sample / dtb, "converged_per_s": float(rb["ok"].sum()) / dtb,
                                           "seconds": dtb, "threads": threads}
    return out


NOMINAL_FP64_TFLOPS = 37.2  # 148 SMs x 64 DFMA lanes x 2 FLOP x 1.965 GHz (no driver-measured FP64 peak exists)


def sampler_leg(c, count, first_index, reps, peak_flops, warm=True, peer=None, world=1, dist=None, windows=None):
    """One timed leg on the sampler path of constraint `c`: `count` device-generated Seeds-U starting at counter
    `first_index` (ccp_sample_project_batch: seed kernel + pipelined projection launches of 2 M seeds + the completing
    launch), per-seed ok + iteration outputs, CUDA events on the launching stream.  peer: PeerPool — the kernel's fused
    gather of the converged states plus the count exchange are inside the timed region.  Returns the record and the
    tensors a check needs."""
    import ctypes as C

    import torch

    from closed_chain_motion_planner_b200 import _capi

    lib, h = c._lib, c._h
    dev = torch.device("cuda", c.device)
    ok = torch.empty(count, dtype=torch.uint8, device=dev)
    it = torch.empty(count, dtype=torch.int32, device=dev)
    n_ok = torch.zeros(1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    fl_iter, fl_tail = c.algorithmicFlops()
    times, counts = [], None
    launches0 = c.launchCount()
    t_open = None
    for r in range(reps + (1 if warm else 0)):
        if t_open is None and (r > 0 or not warm):
            t_open = time.perf_counter()
        a = _capi.SamplerArgs(rng_seed=0, first_index=first_index, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
        n_ok.zero_()
        if peer is not None:
            dist.barrier()  # nobody still reads the pool of the previous repetition
            peer.attach()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.ccp_sample_project_batch(h, C.byref(a), count, 0, None, ok.data_ptr(), it.data_ptr(), None, n_ok.data_ptr(), st)
        assert rc == 0, lib.ccp_last_error(h)
        if peer is not None:
            counts = peer.exchange_counts(n_ok)
        e1.record()
        torch.cuda.synchronize(dev)
        if peer is not None:
            peer.detach()
        if r > 0 or not warm:
            times.append(e0.elapsed_time(e1))
    if windows is not None and t_open is not None:
        windows.append((t_open, time.perf_counter()))  # continuous load: the clock sampler keeps these samples
    ms = sum(times) / len(times)
    iters = int(it.sum(dtype=torch.int64))
    nk = int(n_ok.item())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    w = torch.tensor([float(count), float(nk), float(iters)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    ms = t.item()
    cnt_all, ok_all, it_all = w.tolist()
    flops = it_all * fl_iter + cnt_all * fl_tail
    rec = {"seeds": int(cnt_all), "seeds_per_gpu": count, "ms": ms, "reps": len(times),
           "projections_per_s": cnt_all / ms * 1e3, "converged_per_s": ok_all / ms * 1e3, "ok_fraction": ok_all / cnt_all,
           "mean_iters": it_all / cnt_all,
           "roofline": {"bound": "fp64", "achieved": flops / world / ms / 1e9, "peak": peak_flops / 1e12, "unit": "TFLOP/s",
                        "frac": flops / world / (ms * 1e-3) / peak_flops,
                        "frac_of_nominal": flops / world / (ms * 1e-3) / (NOMINAL_FP64_TFLOPS * 1e12)},
           "gpu_launches": (c.launchCount() - launches0) // (reps + (1 if warm else 0))}
    return rec, ok, n_ok, counts


def gather_check(peer, ok, n_ok, counts, world, rank, dist, dev):
    """N > 1: every rank's copy of the gathered pool must be identical (a wrapping int64 sum over the bit patterns of
    each source rank's rows, compared across ranks), and the exchanged per-rank counts must equal the ranks' own
    recount of their ok flags."""
    import torch

    cl = [int(v) for v in counts.tolist()]
    sums = torch.stack([peer.pool[r, :cl[r]].view(torch.int64).sum() for r in range(world)])
    all_sums = torch.empty((world, world), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sums, sums)
    recount = torch.tensor([int(ok.sum(dtype=torch.int64))], dtype=torch.int64, device=dev)
    all_recount = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_recount, recount)
    same_pools = bool((all_sums == all_sums[0:1]).all())
    counts_ok = [int(v) for v in all_recount.tolist()] == cl and cl[rank] == int(n_ok.item())
    return {"pools_identical_across_ranks": same_pools, "counts_equal_local_recount": counts_ok, "counts": cl,
            "checksum_rank0_rows": int(all_sums[0, 0].item())}


def aux_kernels(pkg, local, peak_flops):
    """The SURVEY §8(f) kernels beside the projection (N = 1 runs): 100 000 discreteGeodesic edges between projected
    states (ccp_geodesic_kernel), 1 M single-arm IK solves (ccp_ik_kernel) and 100 000 goal-sampler targets with 15
    restarts each (ccp_ik_sample_kernel), each with its FP64 roofline fraction on the frozen counts of csrc/ccp_flops.h
    and the kernel's own iteration counters.  CUDA events on the launching stream, best of 3 after a warm-up."""
    import ctypes as C

    import numpy as np
    import torch

    dev = torch.device("cuda", local)

    def best_ms(fn, reps=3):
        fn()
        torch.cuda.synchronize(dev)
        best, out = 1e30, None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        return best, out

    def roof(flops, ms):
        return {"bound": "fp64", "achieved": flops / ms / 1e9, "peak": peak_flops / 1e12, "unit": "TFLOP/s",
                "frac": flops / (ms * 1e-3) / peak_flops, "frac_of_nominal": flops / (ms * 1e-3) / (NOMINAL_FP64_TFLOPS * 1e12)}

    out = {}
    # ---- discreteGeodesic: edges between projected states of the dumbbell manifold ----
    c = pkg.KinematicChainConstraint.from_config("dumbbell", device=local)
    space = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(14), c)
    # the planner's edges (stefanBiPRM.cpp:278-318): every new vertex against its k = 5 nearest roadmap vertices
    smp = space.allocStateSampler(pool_size=1 << 17, rng_seed=3)
    nv, knn = 20_000, 5
    V = smp.sampleUniformBatch(120_000)
    # the pool's row order is whatever the compaction atomics made it: fix it (sort, then a seeded shuffle) so that every
    # run walks the same edges — the launch is bounded by its longest edge, which otherwise changes from run to run
    V = V[torch.argsort(V[:, 0])]
    V = V[torch.randperm(V.shape[0], generator=torch.Generator().manual_seed(0)).to(V.device)][:nv].contiguous()
    dm = torch.cdist(V, V)
    dm.fill_diagonal_(float("inf"))
    nbr = dm.topk(knn, largest=False).indices
    del dm
    E = nv * knn
    frm, to = V.repeat_interleave(knn, dim=0).contiguous(), V[nbr.reshape(-1)].contiguous()
    ms, r = best_ms(lambda: space.discreteGeodesicBatch(frm, to, max_states=40))
    fl_iter, fl_tail = c.algorithmicFlops()
    iters, nproj = int(r.iters.sum(dtype=torch.int64)), int(r.n_states.sum(dtype=torch.int64))
    out["geodesic"] = {"workload": "100 000 discreteGeodesic edges: 20 000 projected dumbbell states, each to its 5 nearest neighbours "
                                   "(delta 0.25, lambda 2)",
                       "kernel": "ccp_geodesic_kernel", "ms": ms, "edges_per_s": E / ms * 1e3,
                       "projections_per_s": nproj / ms * 1e3, "reached_fraction": float(r.reached.float().mean()),
                       "mean_states_per_edge": nproj / E, "newton_iterations": iters,
                       "roofline": roof(iters * fl_iter + nproj * fl_tail, ms)}
    # ---- pose IK ----
    pm = pkg.PandaModel(device=local)
    lb, ub = pm.getJointLimit().T
    rng = np.random.default_rng(0)
    n = 1_000_000
    q_true = lb + (ub - lb) * rng.uniform(0.08, 0.92, (n, 7))
    qd = torch.from_numpy(q_true).to(dev)
    T = torch.empty((n, 12), dtype=torch.float64, device=dev)
    cc = pm._c
    assert cc._lib.ccp_fk_batch(cc._h, 0, qd.data_ptr(), n, 0, T.data_ptr(), torch.cuda.current_stream(dev).cuda_stream) == 0
    T = T.view(n, 3, 4)
    seeds = torch.from_numpy(np.clip(q_true + 0.4 * rng.standard_normal(q_true.shape), lb, ub)).to(dev)
    a_it, a_tail = C.c_double(), C.c_double()
    cc._lib.ccp_algorithmic_flops_ik(C.byref(a_it), C.byref(a_tail))
    ms, r = best_ms(lambda: pm.ikBatch(T, seeds))
    iters = int(r["iters"].sum(dtype=torch.int64))
    out["ik"] = {"workload": "1 000 000 single-arm pose IK solves, seed 0.4 rad (sigma) from a solution, eps 1e-5",
                 "kernel": "ccp_ik_kernel", "ms": ms, "solves_per_s": n / ms * 1e3,
                 "success_fraction": float(r["ok"].float().mean()), "mean_iters": iters / n,
                 "roofline": roof(iters * a_it.value + n * a_tail.value, ms)}
    nt = 100_000
    ms, r = best_ms(lambda: pm.ikSampleBatch(T[:nt].contiguous(), restarts=15, rng_seed=1))
    out["ik_sample"] = {"workload": "100 000 goal-sampler targets x up to 15 restarts (N(mid-range, 0.3) seeds; lowest-numbered success wins, restarts that can no longer win are abandoned)",
                        "kernel": "ccp_ik_sample_kernel", "ms": ms, "targets_per_s": nt / ms * 1e3,
                        "success_fraction": float(r["ok"].float().mean()),
                        "mean_successful_restarts_finished": float(r["n_success"].float().mean())}
    # ---- goal sampling for the whole closed chain (sampleCalibGoal for a batch of object poses) ----
    cs = pkg.KinematicChainConstraint.from_config("stefan", device=local)
    cfg = cs.config
    t_o7 = cs.graspFrames(cfg.t_wo_start, cfg.start)
    ng = 100_000
    To = np.tile(cfg.t_wo_start[None, :3, :], (ng, 1, 1))
    To[:, :, 3] += rng.uniform(-0.06, 0.06, (ng, 3))
    Td = torch.from_numpy(np.ascontiguousarray(To.reshape(ng, 12))).to(dev)
    qref = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(cfg.start, (ng, 14)))).to(dev)
    qo = torch.zeros((ng, 14), dtype=torch.float64, device=dev)
    okg = torch.zeros(ng, dtype=torch.uint8, device=dev)
    to7 = np.ascontiguousarray(t_o7.reshape(2, 12))
    stg = torch.cuda.current_stream(dev).cuda_stream

    def goal():
        assert cs._lib.ccp_goal_sample_batch(cs._h, Td.data_ptr(), ng, to7.ctypes.data, qref.data_ptr(), 15, 5, 0.3, None,
                                             qo.data_ptr(), okg.data_ptr(), stg) == 0
        return okg

    ms, r = best_ms(goal)
    out["goal_sample"] = {"workload": "100 000 object poses around the stefan start pose: both arms' IK (15 restarts each, seeded with "
                                      "the start configuration) -> closed-chain goal configurations (sampleCalibGoal, batched)",
                          "kernels": "ccp_goal_targets_kernel + 2 x ccp_ik_sample_kernel + ccp_goal_combine_kernel", "ms": ms,
                          "poses_per_s": ng / ms * 1e3, "success_fraction": float(r.float().mean())}
    return out


def extra_configs(args, pkg, rank, world, local, dist, peak_flops, windows):
    """BASELINE configs[0], [2], [3] measured in the same process, after the headline region:
      C1  stefan, 10 000 Seeds-U on the reference-faithful CPU oracle (all host cores), the GPU on the same seeds beside it
      C3  Wine_Bottle (joint limits bite: its start sits 7.5e-3 rad from one), 10 M seeds: strong (10 M sharded over
          the N ranks) and weak (10 M per GPU); at N > 1 the fused gather + count exchange inside, with a gather check
      C4  stefan three-arm (21 DoF), 1e3 .. 1e8 seeds on one GPU (N = 1 runs only)"""
    import numpy as np
    import torch

    from closed_chain_motion_planner_b200.dist import PeerPool, gather_capacity, shard_range

    dev = torch.device("cuda", local)
    out = {}
    # ---- C3 ----
    cw = pkg.KinematicChainConstraint.from_config("Wine_Bottle", device=local)
    total = args.c3_seeds
    c3 = {"workload": f"Wine_Bottle (configs/Wine_Bottle.yaml) projection with joint limits, {total} Seeds-U, sampler path "
                      "(device-generated seeds, ok + iteration outputs)"}
    for mode in ("strong", "weak"):
        first, count = shard_range(total, rank, world) if mode == "strong" else (rank * total, total)
        peer = None
        if world > 1:
            try:
                peer = PeerPool(cw, gather_capacity(max(shard_range(total, 0, world)[1], 1) if mode == "strong" else total))
            except Exception as e:  # noqa: BLE001 - every rank fails alike
                c3["peer_memory_error"] = repr(e)
        # the weak leg also feeds the clock record: >= ~1.2 s of continuous load (fixed repetition count, the same on
        # every rank: the fused exchange is a collective)
        reps = 3 if mode == "strong" else max(3, min(60, int(1200.0 / (0.0021 * total / 1000.0)) + 1))
        rec, ok, n_ok, counts = sampler_leg(cw, count, first, reps, peak_flops, peer=peer, world=world, dist=dist,
                                            windows=windows if mode == "weak" else None)
        rec["scaling"] = mode
        if peer is not None:
            rec["exchange"] = ("fused peer stores (%s) + count exchange inside the timed region"
                               % ("one multicast store per 16 B, NVSwitch fan-out" if peer.multicast_ptr else "16-byte stores per peer"))
            rec["gather_check"] = gather_check(peer, ok, n_ok, counts, world, rank, dist, dev)
            assert rec["gather_check"]["pools_identical_across_ranks"] and rec["gather_check"]["counts_equal_local_recount"]
        c3[mode] = rec
        del peer, ok

    out["C3"] = c3
    del cw
    if world > 1:
        return out
    # ---- C4 ----
    c4 = {"workload": "stefan three-arm closed chain (configs/stefan_three_arm.yaml, 21 DoF, co-dimension 4), Seeds-U, "
                      "sampler path, one GPU", "sweep": []}
    ck = pkg.KinematicChainConstraint.from_config("stefan_three_arm", device=local)
    for e in range(3, args.c4_max_exp + 1):
        cnt = 10 ** e
        rec, _, _, _ = sampler_leg(ck, cnt, 0, 3 if e <= 6 else (2 if e == 7 else 1), peak_flops, warm=e <= 7)
        c4["sweep"].append(rec)
    out["C4"] = c4
    # ---- C1 ----
    from closed_chain_motion_planner_b200 import grasping_point
    from oracle.oracle import OracleA

    cs = pkg.KinematicChainConstraint.from_config("stefan", device=local)
    cfg = grasping_point().loadConfig("stefan")
    A = OracleA(cfg.arm_indices)
    A.set_initial_position(cfg.start)
    seeds = A.seeds_uniform(0, 0, 10_000)
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    ra = A.project(seeds, fd=True, nthreads=threads)
    dt = time.perf_counter() - t0
    cs.projectBatch(seeds)  # warm the host path
    t0 = time.perf_counter()
    rg = cs.projectBatch(seeds)
    dg = time.perf_counter() - t0
    okm = rg.ok == 1
    f = A.function(rg.x[okm], nthreads=threads)
    out["C1"] = {"workload": "stefan (configs/stefan.yaml) dual-arm closed-chain projection of 10 000 Seeds-U",
                 "cpu": {"kind": "port", "impl": "oracle A (reference-faithful FD Jacobian + SVD solve)", "cores": threads,
                         "seconds": dt, "projections_per_s": 10_000 / dt, "converged_per_s": float(ra["ok"].sum()) / dt,
                         "ok_fraction": float(ra["ok"].mean()), "mean_iters": float(ra["iters"].mean())},
                 "gpu_host_call": {"api": "ccp_project_batch_host (pageable numpy in, all outputs back)", "seconds": dg,
                                   "projections_per_s": 10_000 / dg, "converged_per_s": float(rg.ok.sum()) / dg},
                 "parity": {"ok_flag_agreement": float(np.mean(rg.ok == ra["ok"])),
                            "converged_flag_agreement": float(np.mean(rg.converged == ra["converged"])),
                            "max_residual_of_ok_states_by_oracle": [float(f[:, 0].max()), float(f[:, 1].max())]}}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = oracle A on all host cores."""
    rank = _env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    per_step = args.ref_sample
    times = []
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_baseline(args.config, per_step, threads)
        if i >= args.warmup:
            times.append(res["seconds"])
    tot = sum(times)
    value = res["ok_fraction"] * per_step * len(times) / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config} closed-chain projection, Seeds-U (bounded sample of the 1M-seed workload)",
                   "seeds_per_step": per_step, "arms": 2},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"first {per_step} Seeds-U of {args.config} per step, oracle A (FD Jacobian + SVD solve, "
                                   "OpenMP over all host cores); the reference itself is not buildable here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "projections_per_s": per_step * len(times) / tot, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="dumbbell")
    ap.add_argument("--count", type=int, default=1_000_000, help="seeds per rank per step")
    ap.add_argument("--cpu-sample", type=int, default=100000, help="seeds of the cpu_baseline leg")
    ap.add_argument("--ref-sample", type=int, default=20000, help="seeds per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--layout", default="aos", choices=["aos", "soa"])
    ap.add_argument("--nccl-gather", action="store_true",
                    help="N > 1: collect the converged states with NCCL all-gathers instead of the kernel's fused peer stores")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 records (BASELINE configs[0], [2], [3])")
    ap.add_argument("--c3-seeds", type=int, default=10_000_000)
    ap.add_argument("--c4-max-exp", type=int, default=8, help="the 21-DoF sweep runs 1e3 .. 1e<this> seeds")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="every launch runs its own stragglers to completion (ccp_project_batch) instead of parking them "
                         "for the next launch (ccp_project_batch_pipelined + one ccp_project_flush)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    import closed_chain_motion_planner_b200 as pkg
    from closed_chain_motion_planner_b200 import _capi
    import ctypes as C

    rank, world, local = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    host_info = bind_to_gpu_numa(torch, local)
    c = pkg.KinematicChainConstraint.from_config(args.config, device=local)
    lib, h = c._lib, c._h
    n, m = c.getAmbientDimension(), c.getCoDimension()
    count = args.count
    layout = pkg.CCP_LAYOUT_AOS if args.layout == "aos" else pkg.CCP_LAYOUT_SOA
    shape = (count, n) if layout == pkg.CCP_LAYOUT_AOS else (n, count)
    stream = torch.cuda.current_stream().cuda_stream

    # measured FP64 peak of this GPU (register-only DFMA chains), before the timed region; the clock sampler runs from
    # here on and its samples are kept per window of continuous load (probe / headline region / C3 weak leg)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    c.fp64PeakProbe(2)
    t_probe = time.perf_counter()
    peak_flops, _ = c.fp64PeakProbe(40)
    probe_window = [(t_probe, time.perf_counter())]
    load_windows = []

    # synthetic seeds, resident in HBM: a different slice of the counter stream per step and per rank
    n_batches = min(args.warmup + args.steps, 8)
    batches = []
    for b in range(n_batches):
        t = torch.empty(shape, dtype=torch.float64, device=dev)
        a = _capi.SamplerArgs(rng_seed=0, first_index=(rank * 64 + b) * count, mode=0, wrap_bounds=0, distance=0.0,
                              near_host=None)
        assert lib.ccp_generate_seeds(h, C.byref(a), count, layout, t.data_ptr(), stream) == 0
        batches.append(t)
    # Outputs.  Steps are PIPELINED (ccp_project_batch_pipelined): the <= ~1e5 samples still iterating when a
    # step's seed list runs dry are carried into the next launch instead of idling the GPU, so a step's per-seed
    # outputs are complete one launch later and every step in flight needs its own arrays (ring of R sets); one
    # ccp_project_flush inside the timed region completes the last step.  --no-pipeline: every launch runs its
    # own stragglers to completion (one output set would do; the ring is kept for symmetry).
    pipelined = not args.no_pipeline
    R = 3
    outs = []
    for r in range(R):
        outs.append(dict(x=torch.empty(shape, dtype=torch.float64, device=dev),
                         ok=torch.empty(count, dtype=torch.uint8, device=dev),
                         cv=torch.empty(count, dtype=torch.uint8, device=dev),
                         it=torch.empty(count, dtype=torch.int32, device=dev),
                         compact=torch.empty((count, n), dtype=torch.float64, device=dev),
                         n_ok=torch.zeros(1, dtype=torch.int64, device=dev)))
    from closed_chain_motion_planner_b200.dist import PeerPool, gather_capacity, gather_converged

    # The path's one exchange step: every rank collects the converged states of all ranks.
    #   fused (default): the projection kernel's epilogue stores each converged state straight into every rank's
    #     pool in symmetric memory (NVLink P2P stores, ccp_set_gather_peers); the 8-byte counts follow the same way
    #     (ccp_publish_count) and a device-side barrier of the symmetric-memory group orders them for the readers;
    #   --nccl-gather / no peer memory: NCCL all-gather of the counts and of the padded compacted states.
    cap = gather_capacity(count)
    peer_pools = None
    exchange_kind = "none"
    if world > 1:
        exchange_kind = "nccl"
        if not args.nccl_gather:
            try:
                peer_pools = [PeerPool(c, cap) for _ in range(R)]
                exchange_kind = "fused"
            except Exception as e:  # every rank fails alike (no P2P on the node)
                if rank == 0:
                    print(f"peer memory unavailable ({e!r}); NCCL gather", file=sys.stderr)
    pool = torch.empty((world, cap, n), dtype=torch.float64, device=dev) if exchange_kind == "nccl" else None
    counts_all = torch.zeros(world, dtype=torch.int64, device=dev)
    max_counts = torch.zeros(world, dtype=torch.int64, device=dev)
    project = lib.ccp_project_batch_pipelined if pipelined else lib.ccp_project_batch

    def pre_launch(i):
        if exchange_kind == "fused":
            peer_pools[i % R].attach()

    # The exchange of step s does not feed step s + 1, so it runs on a side stream behind an event: the ranks' skew
    # at its barrier / collective is not waited for by the next projection launch.  A ring slot is reused only after
    # its exchange has finished (event), and the timed region ends after the side stream has drained.
    xstream = torch.cuda.Stream(device=dev) if world > 1 else None
    xdone = [None] * R

    def exchange(i, o):
        if exchange_kind == "none":
            return
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(xstream):
            xstream.wait_event(ready)
            if exchange_kind == "fused":
                cts = peer_pools[i % R].exchange_counts(o["n_ok"])
                torch.maximum(max_counts, cts, out=max_counts)
            else:
                gather_converged(o["compact"], o["n_ok"], cap, None, pool, counts_all)
                torch.maximum(max_counts, counts_all, out=max_counts)
            xdone[i % R] = torch.cuda.Event()
            xdone[i % R].record()

    def wait_slot(i):
        if xdone[i % R] is not None:
            torch.cuda.current_stream().wait_event(xdone[i % R])

    def step(i, ev0=None, ev1=None):
        o = outs[i % R]
        wait_slot(i)
        o["n_ok"].zero_()
        pre_launch(i)
        if ev0 is not None:
            ev0.record()
        rc = project(h, batches[i % n_batches].data_ptr(), count, layout, o["x"].data_ptr(), o["ok"].data_ptr(),
                     o["cv"].data_ptr(), o["it"].data_ptr(), None,
                     None if exchange_kind == "fused" else o["compact"].data_ptr(), o["n_ok"].data_ptr(), stream)
        assert rc == 0, lib.ccp_last_error(h)
        if ev1 is not None:
            ev1.record()
        exchange(i, o)
        return o

    def flush(i, ev0=None, ev1=None):
        o = outs[i % R]
        wait_slot(i)
        o["n_ok"].zero_()
        pre_launch(i)
        if ev0 is not None:
            ev0.record()
        rc = lib.ccp_project_flush(h, None if exchange_kind == "fused" else o["compact"].data_ptr(), o["n_ok"].data_ptr(), stream)
        assert rc == 0, lib.ccp_last_error(h)
        if ev1 is not None:
            ev1.record()
        exchange(i, o)
        return o

    for i in range(args.warmup):
        o = step(i)
        _ = (o["n_ok"].clone(), o["it"].sum(dtype=torch.int64))  # same (torch) bookkeeping ops as the timed loop: loads them once
    flush(args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    launches0 = c.launchCount()
    t_region = time.perf_counter()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ok = 0
    total_flops = 0.0
    total_iters = 0
    fl_iter, fl_tail = c.algorithmicFlops()
    t_begin.record()
    ok_counts, iter_sums = [], []
    prev = None
    for s in range(args.steps):
        o = step(args.warmup + s, *evs[s])
        # per-step result reads, on device: the states that finished in this launch (8 bytes) and the iteration
        # total of the step whose outputs this launch completed (pipelined: the previous step's)
        ok_counts.append(o["n_ok"].clone())
        done = prev if pipelined else o
        if done is not None:
            iter_sums.append(done["it"].sum(dtype=torch.int64))
        prev = o
    if pipelined:
        o = flush(args.warmup + args.steps, *evs[args.steps])
        ok_counts.append(o["n_ok"].clone())
        iter_sums.append(prev["it"].sum(dtype=torch.int64))
    if xstream is not None:
        torch.cuda.current_stream().wait_stream(xstream)  # the exchanges are part of the timed region
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = c.launchCount() - launches0
    load_windows.append((t_region - 0.05, time.perf_counter()))
    assert int(max_counts.max().item()) <= cap, "gather capacity overflow"
    # N > 1: the pool the LAST full step gathered (its ring slot is untouched since) must be bit-identical on every rank,
    # and the exchanged counts must equal every rank's own count of the states that finished in that launch
    headline_gather_check = None
    if exchange_kind == "fused":
        last = args.warmup + args.steps - 1
        pp, oo = peer_pools[last % R], outs[last % R]
        cl = [int(v) for v in pp.counts.tolist()]
        sums = torch.stack([pp.pool[r, :cl[r]].view(torch.int64).sum() for r in range(world)])
        all_sums = torch.empty((world, world), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_sums, sums)
        mine = torch.tensor([int(oo["n_ok"].item())], dtype=torch.int64, device=dev)
        all_mine = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_mine, mine)
        headline_gather_check = {"pools_identical_across_ranks": bool((all_sums == all_sums[0:1]).all()),
                                 "counts_equal_ranks_own_counts": [int(v) for v in all_mine.tolist()] == cl, "counts": cl}
        assert headline_gather_check["pools_identical_across_ranks"] and headline_gather_check["counts_equal_ranks_own_counts"]

    ms_total = t_begin.elapsed_time(t_end)
    n_timed_launches = args.steps + (1 if pipelined else 0)
    kernel_ms = [a.elapsed_time(b) for a, b in evs[:n_timed_launches]]
    total_ok = sum(int(v.item()) for v in ok_counts)
    total_iters = sum(int(v.item()) for v in iter_sums)
    assert len(iter_sums) == args.steps
    total_flops = total_iters * fl_iter + args.steps * count * fl_tail

    # max over ranks of the timed region; sums over ranks of the work
    tt = torch.tensor([ms_total, sum(kernel_ms)], dtype=torch.float64, device=dev)
    ww = torch.tensor([float(total_ok), float(total_flops), float(total_iters)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
    ms_total_max, kernel_ms_sum_max = tt.tolist()
    ok_all, flops_all, iters_all = ww.tolist()

    # ---- e2e: the host-buffer C-ABI call, pinned host seeds, copies inside the timed region -------------
    e2e = None
    if not args.no_e2e:
        # two sets of pinned buffers: batch k + 1 is submitted before batch k is waited for (the streaming form of the
        # host entry point), so the next batch's launches finish this batch's stragglers and its copies hide behind them
        hs, hx, hok, hit = [], [], [], []
        for r in range(2):
            t = torch.empty((count, n), dtype=torch.float64).pin_memory()
            src = batches[r % len(batches)]
            t.copy_(src if layout == pkg.CCP_LAYOUT_AOS else src.T.contiguous())
            hs.append(t)
            hx.append(torch.empty((count, n), dtype=torch.float64).pin_memory())
            hok.append(torch.empty(count, dtype=torch.uint8).pin_memory())
            hit.append(torch.empty(count, dtype=torch.int32).pin_memory())
        hok_np = [t.numpy() for t in hok]
        torch.cuda.synchronize()
        e_steps = max(3, min(args.steps, 10))

        def sync_call(r):
            assert lib.ccp_project_batch_host(h, hs[r].data_ptr(), count, hx[r].data_ptr(), hok[r].data_ptr(), None,
                                              hit[r].data_ptr(), None) == 0

        def run_sync(steps):
            ok = 0
            for k in range(steps):
                sync_call(k & 1)
                ok += int(np.count_nonzero(hok_np[k & 1]))  # the step's result, read on the host
            return ok

        def run_streaming(steps):
            ok = 0
            tick = C.c_int64(0)
            pending = []
            for k in range(steps + 1):
                if k < steps:
                    r = k & 1
                    assert lib.ccp_project_batch_host_submit(h, hs[r].data_ptr(), count, hx[r].data_ptr(), hok[r].data_ptr(),
                                                             None, hit[r].data_ptr(), None, C.byref(tick)) == 0
                    pending.append((tick.value, r))
                if len(pending) == 2 or (k == steps and pending):
                    t, r = pending.pop(0)
                    assert lib.ccp_project_batch_host_wait(h, t) == 0
                    ok += int(np.count_nonzero(hok_np[r]))  # the step's result, read on the host
            return ok

        def timed_host(fn):
            fn(2)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            ok = fn(e_steps)
            dt = time.perf_counter() - t0
            te = torch.tensor([dt], dtype=torch.float64, device=dev)
            oe = torch.tensor([float(ok)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
                dist.all_reduce(oe, op=dist.ReduceOp.SUM)
            return oe.item(), te.item()

        # ---- compact-output forms (ccp_host_batch_submit / _wait): only the ok states come back, packed, with their
        # seed indices + the per-seed ok / iteration arrays; (a) host seeds in, (b) sampler arguments in (the seeds are
        # generated on the device: the pool refill of jy_ProjectedStateSampler — no H2D at all)
        ccap = int(count * 0.4) + 64
        cst = [torch.empty((ccap, n), dtype=torch.float64).pin_memory() for _ in range(2)]
        cix = [torch.empty(ccap, dtype=torch.int32).pin_memory() for _ in range(2)]
        d2h_compact = [0]

        def make_compact_runner(seeded):
            def run(steps):
                ok = 0
                tick = C.c_int64(0)
                nk = C.c_int64(0)
                pending = []
                keep = []
                for k in range(steps + 1):
                    if k < steps:
                        r = k & 1
                        b = _capi.HostBatch()
                        b.count = count
                        if seeded:
                            sa = _capi.SamplerArgs(rng_seed=0, first_index=(rank * 64 + (k % 8)) * count, mode=0, wrap_bounds=0,
                                                   distance=0.0, near_host=None)
                            keep.append(sa)
                            b.sampler = C.pointer(sa)
                        else:
                            b.seeds_host = hs[r].data_ptr()
                        b.ok_host = hok[r].data_ptr()
                        b.iters_host = hit[r].data_ptr()
                        b.compact_host = cst[r].data_ptr()
                        b.compact_index_host = cix[r].data_ptr()
                        b.compact_capacity = ccap
                        assert lib.ccp_host_batch_submit(h, C.byref(b), C.byref(tick)) == 0, lib.ccp_last_error(h)
                        pending.append((tick.value, r))
                    if len(pending) == 2 or (k == steps and pending):
                        t, r = pending.pop(0)
                        assert lib.ccp_host_batch_wait(h, t, C.byref(nk)) == 0
                        assert 0 <= nk.value <= ccap
                        ok += nk.value  # the step's result: the number of packed rows (checked against the flags below)
                        d2h_compact[0] = count * 5 + nk.value * (n * 8 + 4) + 8
                return ok
            return run

        ok_sync, t_sync = timed_host(run_sync)
        ok_str, t_str = timed_host(run_streaming)
        ok_cmp, t_cmp = timed_host(make_compact_runner(False))
        assert int(np.count_nonzero(hok_np[(e_steps - 1) & 1])) > 0
        d2h_cmp = d2h_compact[0]
        ok_sed, t_sed = timed_host(make_compact_runner(True))
        d2h_sed = d2h_compact[0]

        # ---- host ceiling: nothing but the copies of the full-output form, on every rank at once (what the box's host
        # side can move through pinned memory; the full-output e2e cannot beat count / this time)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        dstage = torch.empty((count, n), dtype=torch.float64, device=dev)
        dflags = torch.empty(count * 5, dtype=torch.uint8, device=dev)
        hflags = torch.empty(count * 5, dtype=torch.uint8).pin_memory()

        def run_copies(steps):
            for k in range(steps):
                with torch.cuda.stream(s_in):
                    dstage.copy_(hs[k & 1], non_blocking=True)
                with torch.cuda.stream(s_out):
                    hx[k & 1].copy_(dstage, non_blocking=True)
                    hflags.copy_(dflags, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
            return 0

        _, t_cpy = timed_host(run_copies)
        bytes_step = count * n * 8 + count * (n * 8 + 5)
        ceiling_pps = world * count * e_steps / t_cpy
        e2e = {"value": ok_str / t_str, "unit": UNIT, "h2d_bytes_per_step": count * n * 8,
               "d2h_bytes_per_step": count * (n * 8 + 1 + 4), "steps": e_steps,
               "projections_per_s": world * count * e_steps / t_str,
               "host_ceiling": {"what": "the same pinned H2D (states) + D2H (states, ok, iters) copies alone, all ranks at once, "
                                        "two streams per rank, no kernels",
                                "aggregate_gb_per_s": world * bytes_step * e_steps / t_cpy / 1e9,
                                "projections_per_s": ceiling_pps,
                                "device_rate_projections_per_s": world * count * args.steps / (ms_total_max * 1e-3),
                                # the full-output e2e is bounded by the slower of the two: the host's copies, the kernels
                                "e2e_fraction_of_bound": (world * count * e_steps / t_str)
                                / min(ceiling_pps, world * count * args.steps / (ms_total_max * 1e-3))},
               "compact_outputs": {"value": ok_cmp / t_cmp, "projections_per_s": world * count * e_steps / t_cmp,
                                   "h2d_bytes_per_step": count * n * 8, "d2h_bytes_per_step": d2h_cmp,
                                   "api": "ccp_host_batch_submit / _wait: pinned host states in; ok + iters per seed and the ok "
                                          "states packed (with seed indices) out"},
               "seeded": {"value": ok_sed / t_sed, "projections_per_s": world * count * e_steps / t_sed,
                          "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h_sed,
                          "api": "ccp_host_batch_submit / _wait with sampler arguments (the batched jy_ProjectedStateSampler "
                                 "refill): seeds generated on the device from the counter stream, packed ok states + ok/iters "
                                 "out; not an e2e number in the contract's sense (no inputs cross PCIe), reported beside it"},
               "host": host_info,
               "synchronous_call": {"value": ok_sync / t_sync, "projections_per_s": world * count * e_steps / t_sync,
                                    "api": "ccp_project_batch_host, one blocking call per step"},
               "api": "ccp_project_batch_host_submit / _wait, two batches in flight: pinned host AOS states in, states + ok "
                      "+ iters out; chunked H2D | pipelined projection launches | D2H of completed chunks on three streams; "
                      "every step's submit, wait and host-side read of its ok flags inside the timed region"}

    configs = None
    if not args.no_configs:
        configs = extra_configs(args, pkg, rank, world, local, dist if world > 1 else None, peak_flops, load_windows)
    clocks = None
    if rank == 0:
        clocks = sampler.stop(load_windows)
        clocks["windows"] = "headline timed region + the C3 weak leg (continuous load)"
        clocks["fp64_probe"] = ClockSampler.summarise(sampler.lines, probe_window)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    secs = ms_total_max * 1e-3
    value = ok_all / secs
    ksecs = kernel_ms_sum_max * 1e-3
    achieved = (flops_all / world) / ksecs  # per-GPU FLOP/s of the projection kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # dram bytes read + written per launch of the projection kernel, from the committed `ncu --set full` capture
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("config") == args.config and tr.get("count") == count:
            traffic = tr.get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    alg_bytes = (2 * n * 8 + 1 + 1 + 4) * count  # seeds in, states out, ok, converged, iters
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config} (configs/{args.config}.yaml) closed-chain projection, {count} uniform seeds per GPU per step",
                   "seeds_per_gpu_per_step": count, "arms": c.k_, "layout": args.layout,
                   "tolerances": [1e-3, 5e-3], "step": 0.30, "max_iter": 250,
                   "l2": f"inputs+outputs {2 * count * n * 8 / 1e6:.0f} MB per step exceed the 126 MB L2; seed batches rotate over {n_batches} buffers",
                   "launches": ("pipelined: ccp_project_batch_pipelined per step (stragglers carried into the next launch) + one "
                                "ccp_project_flush, all inside the timed region") if pipelined else "ccp_project_batch per step",
                   "exchange": {"none": "none",
                                "fused": "per step: the projection kernel stores every converged state into all ranks' pools "
                                         "(symmetric memory, NVLink P2P stores); the 8-byte counts follow by peer stores + the "
                                         "symmetric-memory group's device-side barrier (no collective-library call in the step)",
                                "nccl": "per step: NCCL all_gather of counts + padded compacted converged states"}[exchange_kind]},
        "projections_per_s": world * count * args.steps / secs,
        "ok_fraction": ok_all / (world * count * args.steps),
        "mean_iters": iters_all / (world * count * args.steps),
        "roofline": {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak_flops / 1e12, "unit": "TFLOP/s",
                     "frac": achieved / peak_flops, "frac_of_nominal": achieved / (NOMINAL_FP64_TFLOPS * 1e12),
                     "nominal_peak": NOMINAL_FP64_TFLOPS, "traffic": traffic,
                     "peak_source": "measured in this run: ccp_fp64_peak_probe (register-only DFMA chains, best of 40; SM clock during "
                                    "the probe in clocks.fp64_probe); MEASURED_PEAKS.json has no FP64 entry; nominal = 148 SM x 64 "
                                    "lanes x 2 x 1.965 GHz",
                     "flops_per_iteration": fl_iter, "flops_per_tail": fl_tail,
                     # second figure: the transcendental calls the headline count leaves out, expanded at
                     # a stated cost of 50 FLOP per sincos and 60 per atan2 (7K sincos + (K-1) atan2 per evaluation)
                     "achieved_with_transcendentals_expanded_tflops":
                         (achieved + (iters_all + world * count * args.steps) / world * (7 * c.k_ * 50 + (c.k_ - 1) * 60) / ksecs) / 1e12,
                     "kernel_ms_per_launch": kernel_ms_sum_max / n_timed_launches, "kernel_launches_timed": n_timed_launches,
                     "hbm": {"achieved_gbs": alg_bytes / (ksecs / args.steps) / 1e9, "peak_gbs": hbm_peak,
                             "frac": alg_bytes / (ksecs / args.steps) / 1e9 / hbm_peak,
                             "algorithmic_bytes_per_projection": 2 * n * 8 + 6,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        "gather_multicast": bool(peer_pools and peer_pools[0].multicast_ptr),
        "gather_check": headline_gather_check,
        "gpu_launches": launches, "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    if configs:
        line["configs"] = configs
    if world == 1 and not args.no_configs:
        line["aux_kernels"] = aux_kernels(pkg, local, peak_flops)
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = cpu_baseline(args.config, args.cpu_sample, threads, with_engine_arithmetic=True)
        line["cpu_baseline"] = {"value": cb["converged_per_s"], "unit": UNIT, "cores": threads, "kind": "port",
                                "projections_per_s": cb["projections_per_s"], "mean_iters": cb["mean_iters"],
                                "single_thread_projections_per_s": cb.get("single_thread_projections_per_s"),
                                "engine_arithmetic_on_cpu": cb.get("engine_arithmetic_on_cpu"),
                                "sample": f"first {args.cpu_sample} Seeds-U of {args.config}, oracle A (reference-faithful FD "
                                          f"Jacobian + SVD solve) on {threads} threads, {cb['seconds']:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
