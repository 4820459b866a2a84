#!/usr/bin/env python
"""bench.py -- collision checks/sec on a synthetic pose batch (BASELINE.json metric).

A step = one pass of the hot path (SE3RigidBodyScenario::isValid(q), reference
include/mpl/demo/se3_rigid_body_scenario.hpp:119-126) over one batch of 2^20 seeded uniformly
sampled poses of the Alpha 1.5 puzzle scene (BASELINE.json configs[1]).

  value  device-resident poses -> verdict bytes, one checkStatesKernel launch per step
  e2e    the same batches through the C-ABI call mplb_check_states with pinned HOST buffers (H2D of the
         poses and D2H of the verdicts inside the timed region), issued by two caller threads the way the
         reference's OpenMP planner threads call isValid (prrt.hpp:140-152): one call's verdict read-back and
         drain overlap the next call's input copy.  `e2e.single_caller` is the same with one caller.
  e2e_device_sampled   the same metric when the poses never cross PCIe: mplb_sample_states_device (the same
         Philox stream) + mplb_check_states_device, only the verdict bytes come back
  roofline   HBM bytes the device's own traversal would need uncached (28 + 1 + 120 n_bv + 72 n_tri per
         check, n_* = the device counters) / kernel time vs the measured HBM peak, next to the physical DRAM
         traffic and the units that actually bind (ncu, profiles/)
  cpu_baseline   the FP64 oracle (oracle/, FCL restatement, OpenMP over all host cores) on the same batch
  extras (N = 1)   the other rows of the path, each next to the CPU path and with its mismatch count: a
         Twistycool 2^20-pose line (the north-star target), edge batches (config 3), nearest neighbour and
         k-NN + connect (config 4), Fetch states and edges (config 5), RRT time-to-solution
  multi_gpu (N > 1)   configs 3-5 sharded over the ranks (edge batches balanced by step count, Fetch
         configurations, Home k-NN + connect) and independent planner shards exchanging their solution
         over NCCL (SolutionExchange), each with the whole-job throughput at this N

`--impl reference` times the CPU path alone on the same 2^20-pose batches (the reference's own FCL path
cannot be built: FCL/libccd/Eigen/nigh/assimp are absent).  Multi-GPU: one process per GPU
under torchrun, BVHs replicated, each rank checks its own 2^20-pose shard of the stream (weak scaling, no
data-path collective); time = max over ranks.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses  # noqa: E402

N_POSES = 1 << 20
ROTATE = 4  # distinct pose batches cycled through: 4 x 58.7 MB > 126 MB L2


def load_scene(name):
    d = np.load(os.path.join(ROOT, "tests", "golden", "scenes", name + ".npz"))
    return d["env"], d["robot"]


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1: ignore it for the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_facts(workload):
    """What the committed ncu capture of the dominant kernel says (profiles/traffic.json): DRAM bytes per launch and the
    utilisation of the units that bind, or {}."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        v = d.get(workload)
        return v if isinstance(v, dict) else ({"dram_bytes": v} if v else {})
    except Exception:
        return {}


def hbm_roofline(kernel, bytes_per_unit, units, ms, note, facts=None):
    """roofline block: algorithmic bytes / kernel time against the measured HBM peak."""
    peak, how = measured_peak()
    achieved = bytes_per_unit * units / (ms * 1e-3) / 1e9
    facts = facts or {}
    out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "traffic": facts.get("dram_bytes"), "peak_source": how, "kernel": kernel, "kernel_ms": ms,
           "algorithmic_bytes_per_unit": bytes_per_unit, "note": note}
    if facts.get("binding_units"):
        out["binding_units"] = facts["binding_units"]
        out["ncu_summary"] = facts.get("summary")
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class HostMemoryNearGpu:
    """`numactl --preferred=<the GPU's NUMA node>` for the allocations made inside the block: the pinned batches of
    rank r are placed in the memory of the socket GPU r hangs off, so that with 8 ranks the input copies do not all
    read one socket's DRAM (and half of them across the socket link).  Best effort: a host that exposes no NUMA
    topology, or a cpuset that does not allow the node, leaves the placement as it was and says so."""
    MPOL_DEFAULT, MPOL_PREFERRED, SYS_SET_MEMPOLICY = 0, 1, 238   # x86_64

    def __init__(self, device, enabled=True):
        self.node, self.note = -1, "default placement"
        if not enabled:
            self.note = "default placement (--no-numa)"
            return
        try:
            import ctypes
            import platform
            if platform.machine() != "x86_64":
                raise OSError("syscall number known for x86_64 only")
            buf = ctypes.create_string_buffer(32)
            rt = ctypes.CDLL("libcudart.so.12")
            if rt.cudaDeviceGetPCIBusId(buf, 32, int(device)) != 0:
                raise OSError("cudaDeviceGetPCIBusId failed")
            with open("/sys/bus/pci/devices/%s/numa_node" % buf.value.decode().lower()) as f:
                self.node = int(f.read())
            if self.node < 0:
                self.note = "default placement (the host reports no NUMA node for the GPU)"
        except (OSError, ValueError) as e:
            self.node, self.note = -1, "default placement (%s)" % e

    def _policy(self, mode, node):
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        if node < 0:
            return libc.syscall(self.SYS_SET_MEMPOLICY, mode, None, 0)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        return libc.syscall(self.SYS_SET_MEMPOLICY, mode, mask, 16 * 64 + 1)

    def __enter__(self):
        if self.node >= 0:
            if self._policy(self.MPOL_PREFERRED, self.node) == 0:
                self.note = "pinned batches preferred on NUMA node %d (the GPU's)" % self.node
            else:
                import ctypes
                self.note = "default placement (set_mempolicy: errno %d)" % ctypes.get_errno()
                self.node = -1
        return self

    def __exit__(self, *exc):
        if self.node >= 0:
            self._policy(self.MPOL_DEFAULT, -1)
        return False


def cpu_baseline(env, robot, poses, threads, target_seconds=12.0):
    """Oracle O2 (FCL traversal order, first-hit exit) on a bounded sample sized for ~10-20 s."""
    from oracle import oracle_py as O
    orc = O.SE3Oracle(env, robot)
    threads = threads or host_threads()
    orc.check_states(poses[:8192], O.BVH, threads=threads)   # warm-up
    n = len(poses)
    ctr = O.Counters()
    reps, dt, valid = 0, 0.0, None
    while dt < target_seconds and reps < 200:      # whole batch, repeated until ~target_seconds of CPU work
        t = time.perf_counter(); valid = orc.check_states(poses, O.BVH, threads=threads, counters=ctr)
        dt += time.perf_counter() - t
        reps += 1
    ctr.bv_tests //= reps; ctr.tri_tests //= reps
    dt /= reps
    one = time.perf_counter(); orc.check_states(poses[:1 << 16], O.BVH, threads=1); one = (1 << 16) / (time.perf_counter() - one)
    return {"value": n / dt, "unit": "checks/s", "cores": int(threads), "kind": "port",
            "sample": "the step's %d-pose batch x %d passes, oracle O2 (FP64 FCL restatement, OpenMP dynamic,256), "
                      "%.3f s per pass" % (n, reps, dt),
            "single_core_value": one,
            "n_bv_per_check": ctr.bv_tests / n, "n_tri_per_check": ctr.tri_tests / n,
            "valid_fraction": float(valid.mean())}, valid, n


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t)
    return min(ts), r


def pose_batch_line(name, local_rank, cpu_seconds):
    """A 2^20-pose batch of another scene measured like the contract line: device resident, through the C ABI, CPU."""
    import torch
    import mplambda_b200 as M
    env, robot = load_scene(name)
    sc = SE3_SCENES[name]
    s = M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], 0.1, device=local_rank)
    host = [se3_poses(N_POSES, sc["min"], sc["max"], SEED, start=r * N_POSES) for r in range(ROTATE)]
    dev = [torch.from_numpy(h).cuda() for h in host]
    pin = [torch.from_numpy(h).pin_memory() for h in host]
    out = torch.empty(N_POSES, dtype=torch.uint8, device="cuda")
    pout = torch.empty(N_POSES, dtype=torch.uint8).pin_memory()
    st = torch.cuda.current_stream()
    for k in range(2 * ROTATE):   # every pinned batch crosses the link before the timing (see main)
        s.check_states_device(dev[k % ROTATE].data_ptr(), N_POSES, out.data_ptr(), st.cuda_stream)
        s.check_states_ptr(pin[k % ROTATE].data_ptr(), N_POSES, pout.data_ptr())
    torch.cuda.synchronize()
    s.stats(reset=True)
    steps = 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for k in range(steps):
        s.check_states_device(dev[k % ROTATE].data_ptr(), N_POSES, out.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cnt = s.stats(reset=True)
    t = time.perf_counter()
    for k in range(steps):
        s.check_states_ptr(pin[k % ROTATE].data_ptr(), N_POSES, pout.data_ptr())
    e2e_ms = (time.perf_counter() - t) / steps * 1e3
    cb, cpu_valid, _ = cpu_baseline(env, robot, host[0], 0, cpu_seconds)
    got = s.check_states(host[0])
    nbv, ntri = cnt["bv_tests"] / (steps * N_POSES), cnt["tri_tests"] / (steps * N_POSES)
    line = {"metric": "collision checks/sec (pose batch)", "workload": "%s, 2^20 uniformly sampled poses per step" % name,
            "value": N_POSES / (ms * 1e-3), "unit": "checks/s", "ms_per_step": ms,
            "e2e": {"value": N_POSES / (e2e_ms * 1e-3), "unit": "checks/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": N_POSES * 56, "d2h_bytes_per_step": N_POSES,
                    "api": "mplb_check_states (pinned host buffers, one caller)"},
            "cpu_baseline": cb, "verdict_mismatches_vs_cpu": int((got != cpu_valid).sum()),
            "device_vs_cpu": N_POSES / (ms * 1e-3) / cb["value"], "e2e_vs_cpu": N_POSES / (e2e_ms * 1e-3) / cb["value"],
            "roofline": hbm_roofline("checkStatesKernel", 29 + 120 * nbv + 72 * ntri, N_POSES, ms,
                                     "device traversal counts (%.1f box / %.2f triangle tests per check); the hierarchy is "
                                     "cache resident, what binds is the SM (see the contract line)" % (nbv, ntri))}
    s.close()
    return line


def extras(local_rank, cpu_seconds):
    """The other rows of the hot path, each timed end to end through the C ABI next to the
    oracle on the host cores: edge batches (config 3), nearest neighbour (config 4), Fetch (config 5).
    Reported under "extras"; the contract metric above is unaffected."""
    import torch
    import mplambda_b200 as M
    from mplambda_b200.workload import FETCH_TASKS, fetch_configs, frame_from_option
    from oracle import oracle_py as O
    thr = host_threads()
    out = {}

    # ---- the north-star target line: Twistycool pose batches (BASELINE.json "Target")
    out["poses_twistycool"] = pose_batch_line("twistycool", local_rank, min(cpu_seconds, 6.0))
    # ---- isValid(from,to) batches: 65,536 seeded edges (valid start, random end), resolution 0.1
    for name in ("cubicles", "apartment"):
        env, robot = load_scene(name)
        sc = SE3_SCENES[name]
        s = M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], 0.1, device=local_rank)
        P = se3_poses(8 << 16, sc["min"], sc["max"], SEED)
        A = P[s.check_states(P)][:1 << 16]
        B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
        s.check_edges(A[:512], B[:512])
        # end to end from page-locked host buffers (like the pose batches' e2e): the end states start crossing the link
        # while the host plans the step counts; and from ordinary (pageable) arrays, through the bounce buffers
        Ap, Bp = torch.from_numpy(A).pin_memory().numpy(), torch.from_numpy(B).pin_memory().numpy()
        s.check_edges(A, B)
        dt_pageable, _ = best(lambda: s.check_edges(A, B))
        s.check_edges(Ap, Bp)
        s.stats(reset=True)
        dt, (v, checked) = best(lambda: s.check_edges(Ap, Bp, return_states_checked=True))
        cnt = s.stats()
        steps = s.edge_steps(A, B)
        # device resident: the three passes of checkEdgesKernel alone
        dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        dS = torch.from_numpy(steps.astype(np.int32)).cuda()
        dV = torch.empty(len(A), dtype=torch.uint8, device="cuda")
        cs = torch.cuda.current_stream()
        for _ in range(2):
            s.check_edges_device(dA.data_ptr(), dB.data_ptr(), dS.data_ptr(), len(A), int(steps.max()), dV.data_ptr(), cs.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cs)
        for _ in range(5):
            s.check_edges_device(dA.data_ptr(), dB.data_ptr(), dS.data_ptr(), len(A), int(steps.max()), dV.data_ptr(), cs.cuda_stream)
        e1.record(cs)
        torch.cuda.synchronize()
        kms = e0.elapsed_time(e1) / 5
        orc = O.SE3Oracle(env, robot, check_resolution=0.1)
        m = 4096
        t = time.perf_counter(); w = orc.check_edges(A[:m], B[:m], O.BVH, threads=thr); dc = time.perf_counter() - t
        per_state = 29 + (120 * cnt["bv_tests"] + 72 * cnt["tri_tests"]) / max(1, 3 * checked)
        out["edges_" + name] = {
            "metric": "edge validity checks/sec (isValid(from,to), resolution 0.1)", "value": len(A) / dt, "unit": "edges/s",
            "edges": int(len(A)), "interpolated_states": int(steps.sum()), "device_state_evaluations": int(checked),
            "valid_fraction": float(v.mean()), "ms": dt * 1e3, "api": "mplb_check_edges (pinned host buffers)",
            "pageable_inputs_ms": dt_pageable * 1e3, "device_resident_ms": kms,
            "device_resident_value": len(A) / (kms * 1e-3),
            "cpu_baseline": {"value": m / dc, "unit": "edges/s", "cores": thr, "kind": "port", "sample": "first %d edges" % m},
            "verdict_mismatches_vs_cpu": int((w != v[:m]).sum()),
            "roofline": hbm_roofline("checkEdgesKernel (3 passes) + edgeConstKernel", per_state + 0.0, int(checked), kms,
                                     "bytes of the %d states the device evaluates (%.2f %% of the interpolated states: the rest "
                                     "is certified per interval); bound by instruction fetch and latency, not by HBM "
                                     "(profiles/r1_edges_cubicles_v4_ncu_summary.txt)" % (checked, 100.0 * checked / steps.sum()))}
        s.close()
    # ---- nearest neighbour: 65,536 queries against 16,384 keys (SE3 metric), 1-NN and k-NN (k = 32)
    sc = SE3_SCENES["home"]
    keys = se3_poses(1 << 14, sc["min"], sc["max"], SEED)
    qs = se3_poses(1 << 16, sc["min"], sc["max"], SEED + 2)
    nn = M.Nigh(M.nn.SE3, device=local_rank)
    nn.set_index(False)   # this row is the brute-force scan (16,384 keys is where the index takes over by default)
    nn.insert(keys)
    nn.nearest(qs[:256])
    dt, (idx, dist) = best(lambda: nn.nearest(qs))
    dtk, _ = best(lambda: nn.knearest(qs[:1 << 14], 32), reps=2)
    dq = torch.from_numpy(qs).cuda()
    di = torch.empty(len(qs), dtype=torch.int64, device="cuda")
    dd = torch.empty(len(qs), dtype=torch.float64, device="cuda")
    cs = torch.cuda.current_stream()
    nn.nearest_device(dq.data_ptr(), len(qs), di.data_ptr(), dd.data_ptr(), cs.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cs)
    for _ in range(5):
        nn.nearest_device(dq.data_ptr(), len(qs), di.data_ptr(), dd.data_ptr(), cs.cuda_stream)
    e1.record(cs)
    torch.cuda.synchronize()
    nms = e0.elapsed_time(e1) / 5
    m = 2048
    t = time.perf_counter(); widx, wdist = O.nn_nearest(keys, qs[:m], threads=thr); dc = time.perf_counter() - t
    out["nn_home"] = {"metric": "nearest-neighbour queries/sec (brute-force scan, SE3 metric, index off)", "value": len(qs) / dt,
                      "unit": "queries/s", "keys": int(len(keys)), "queries": int(len(qs)), "ms": dt * 1e3,
                      "device_resident_ms": nms, "knn32_queries_per_s": (1 << 14) / dtk,
                      "cpu_baseline": {"value": m / dc, "unit": "queries/s", "cores": thr, "kind": "port",
                                       "sample": "first %d queries, brute force" % m},
                      "index_mismatches_vs_cpu": int((widx != idx[:m]).sum()),
                      "distance_mismatches_vs_cpu": int((wdist != dist[:m]).sum()),
                      "roofline": hbm_roofline("nnBound32Kernel + nnThresholdKernel + nnRefineKernel + nnMergeKernel",
                                               28.0 * len(keys) / 128 + 56 + 16, len(qs), nms,
                                               "key tiles are staged in shared memory and shared by the 128 queries of a CTA: n x "
                                               "28 B / 128 of key reads per query + the query in + (index, distance) out; the scan "
                                               "is bound by FP32 issue (%.0f G pair distances/s), not by HBM"
                                               % (len(qs) * len(keys) / (nms * 1e-3) / 1e9))}
    # ---- beyond brute force (nigh's KD-tree strategy, prrt.hpp:35-42): 10^6 keys, through the grid index
    #      and with the index off; k = 45 is PCForest's k at that tree size (pcforest.hpp:512-520)
    big = se3_poses(1000000, sc["min"], sc["max"], SEED + 5)
    q1k = qs[:1024]
    dq1 = dq[:1024]
    kk = 45
    dik = torch.empty((1024, kk), dtype=torch.int64, device="cuda")
    ddk = torch.empty((1024, kk), dtype=torch.float64, device="cuda")
    row = {}
    for label, on in (("index", True), ("brute_force", False)):
        nb = M.Nigh(M.nn.SE3, device=local_rank)
        nb.set_index(on)
        nb.insert(big)
        nb.knearest_device(dq1.data_ptr(), 1024, kk, dik.data_ptr(), ddk.data_ptr(), cs.cuda_stream)
        nb.nearest_device(dq.data_ptr(), len(qs), di.data_ptr(), dd.data_ptr(), cs.cuda_stream)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(cs)
        for _ in range(3):
            nb.knearest_device(dq1.data_ptr(), 1024, kk, dik.data_ptr(), ddk.data_ptr(), cs.cuda_stream)
        ev[1].record(cs)
        nb.nearest_device(dq.data_ptr(), len(qs), di.data_ptr(), dd.data_ptr(), cs.cuda_stream)
        ev[2].record(cs)
        torch.cuda.synchronize()
        row[label] = {"knn45_1024_queries_ms": ev[0].elapsed_time(ev[1]) / 3, "nearest_65536_queries_ms": ev[1].elapsed_time(ev[2]),
                      "idx": dik.cpu().numpy().copy(), "idx1": di.cpu().numpy().copy(), "info": nb.index_info()}
        nb.close()
    m = 32
    t = time.perf_counter(); wi, _ = O.nn_knearest(big, q1k[:m], kk, threads=thr); dc = time.perf_counter() - t
    ims = row["index"]["knn45_1024_queries_ms"]
    out["nn_index_1m"] = {
        "metric": "k-NN queries/sec against 10^6 SE(3) keys through the grid index (k = 45, 1,024 queries per call, device resident)",
        "value": 1024 / (ims * 1e-3), "unit": "queries/s", "ms": ims, "brute_force_ms": row["brute_force"]["knn45_1024_queries_ms"],
        "nearest_65536_queries_ms": row["index"]["nearest_65536_queries_ms"],
        "nearest_65536_queries_brute_force_ms": row["brute_force"]["nearest_65536_queries_ms"],
        "cells": int(row["index"]["info"]["cells"]), "slow_queries": int(row["index"]["info"]["slow_queries"]),
        "equal_to_brute_force": bool((row["index"]["idx"] == row["brute_force"]["idx"]).all()
                                     and (row["index"]["idx1"] == row["brute_force"]["idx1"]).all()),
        "cpu_baseline": {"value": m / dc, "unit": "queries/s", "cores": thr, "kind": "port", "sample": "first %d queries, brute force" % m},
        "neighbour_mismatches_vs_cpu": int((wi != row["index"]["idx"][:m]).sum()),
        "roofline": hbm_roofline("nnGridSearchKernel", 32.0 * 2 * 4000 + 56 + 16 * kk, 1024, ims,
                                 "per query: the FP32 keys (32 B) of the ~4,000 keys in reach, read twice (threshold pass, "
                                 "candidate pass), the query in and k (index, distance) pairs out; one warp per query, bound by "
                                 "the latency of its dependent row scans, not by HBM")}
    del big
    # ---- config 4: the addNodeNear pattern batched (pcforest.hpp:512-566) on Home: k-NN of each new node in the
    #      tree, then the motion from the node to each of its k neighbours.  Only indices move: the new nodes go up
    #      once (56 B each), the neighbours are named by index into the tree's device-resident keys.
    env, robot = load_scene("home")
    s = M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], 0.1, device=local_rank)
    tree = keys[s.check_states(keys)]
    new = qs[s.check_states(qs)][:4096]
    k = int(math.ceil(math.e * (7.0 / 6.0) * math.log(len(tree) + 1)))
    nn2 = M.Nigh(M.nn.SE3, device=local_rank)
    nn2.insert(tree)
    h_new = torch.from_numpy(new).pin_memory()
    d_new = torch.empty((len(new), 7), dtype=torch.float64, device="cuda")
    d_i = torch.empty((len(new), k), dtype=torch.int64, device="cuda")
    d_d = torch.empty((len(new), k), dtype=torch.float64, device="cuda")
    d_v = torch.empty(len(new) * k, dtype=torch.uint8, device="cuda")
    h_i = torch.empty((len(new), k), dtype=torch.int64).pin_memory()
    h_v = torch.empty(len(new) * k, dtype=torch.uint8).pin_memory()

    def near_and_connect():
        d_new.copy_(h_new, non_blocking=True)
        nn2.knearest_device(d_new.data_ptr(), len(new), k, d_i.data_ptr(), d_d.data_ptr(), cs.cuda_stream)
        s.check_edges_indexed_device(d_new.data_ptr(), len(new), None, nn2.keys_device(), nn2.size(), d_i.data_ptr(),
                                     len(new) * k, d_v.data_ptr(), from_div=k, stream=cs.cuda_stream)
        h_i.copy_(d_i, non_blocking=True)
        h_v.copy_(d_v, non_blocking=True)
        cs.synchronize()
        return h_i.numpy(), h_v.numpy().astype(bool)

    def near_and_connect_host():   # round 1's path: neighbours to the host, both end states of every motion back up
        ki_, _ = nn2.knearest(new, k)
        return ki_, s.check_edges(np.repeat(new, k, axis=0), tree[ki_.reshape(-1)])
    near_and_connect()
    s.stats(reset=True)
    dt, (ki, ev) = best(near_and_connect, reps=3)
    amb = s.stats()["ambiguous_steps"]
    near_and_connect_host()
    dth, (kih, evh) = best(near_and_connect_host, reps=2)
    m = 48
    orc = O.SE3Oracle(env, robot, check_resolution=0.1)
    t = time.perf_counter()
    wi, _ = O.nn_knearest(tree, new[:m], k, threads=thr)
    wv = orc.check_edges(np.repeat(new[:m], k, axis=0), tree[wi.reshape(-1)], O.BVH, threads=thr)
    dc = time.perf_counter() - t
    out["knn_connect_home"] = {
        "metric": "new nodes/sec through k-NN + k motion checks (PCForest addNodeNear pattern)", "value": len(new) / dt,
        "unit": "nodes/s", "tree_nodes": int(len(tree)), "new_nodes": int(len(new)), "k": k, "edges": int(len(new) * k),
        "edge_valid_fraction": float(ev.mean()), "ms": dt * 1e3,
        "api": "mplb_nn_knearest_device + mplb_check_edges_indexed_device; H2D 56 B per new node, D2H k x 9 B per node",
        "host_staged_ms": dth * 1e3, "host_staged_value": len(new) / dth,
        "ambiguous_step_counts": int(amb), "verdicts_equal_host_staged_path": bool((ev == evh).all() and (ki == kih).all()),
        "cpu_baseline": {"value": m / dc, "unit": "nodes/s", "cores": thr, "kind": "port", "sample": "first %d new nodes" % m},
        "neighbour_mismatches_vs_cpu": int((wi != ki[:m]).sum()), "verdict_mismatches_vs_cpu": int((wv != ev[:m * k]).sum())}
    s.close()
    # ---- Fetch: 2^18 random configurations, AUTOLAB environment, task-001 frame
    env = np.load(os.path.join(ROOT, "tests", "golden", "scenes", "autolab.npz"))["env"]
    frame = frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"])
    f = M.FetchScenario(frame, env, checkResolution=0.01, device=local_rank)
    Q = fetch_configs(1 << 18, SEED)
    f.check_states(Q[:1024])
    f.check_states(Q)
    dt, v = best(lambda: f.check_states(Q))
    m = 1 << 15
    t = time.perf_counter(); w = O.FetchOracle(env, frame, 0.01).check_states(Q[:m], threads=thr); dc = time.perf_counter() - t
    out["fetch_states"] = {"metric": "Fetch validity checks/sec (FK + 24 MPR pairs + 12 shapes vs mesh)", "value": len(Q) / dt,
                           "unit": "checks/s", "configs": int(len(Q)), "valid_fraction": float(v.mean()), "ms": dt * 1e3,
                           "cpu_baseline": {"value": m / dc, "unit": "checks/s", "cores": thr, "kind": "port",
                                            "sample": "first %d configurations" % m},
                           "verdict_mismatches_vs_cpu": int((w != v[:m]).sum()),
                           "roofline": hbm_roofline("fetchFkKernel .. fetchLeafKernel (through mplb_check_states)",
                                                    64 + 1 + 12 * 96, len(Q), dt * 1e3,
                                                    "64 B in, 1 B out and the 12 link frames (96 B each) written and read once per "
                                                    "configuration; the time is end to end (copies included); the kernels are FP64 "
                                                    "MPR at 9 of 32 lanes (profiles/r1_fetch_states_ncu_summary.txt), nowhere near HBM")}
    # ---- the metric's second half: RRT time-to-solution, the same batch-issuing planner (mplambda_b200/planner.py) with the
    #      GPU scenario + GPU nearest neighbours and with the CPU path (oracle on all host cores) as its backend; and the
    #      device-resident loop (mplb_rrt_*)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cpu_backend import CpuNN, CpuScenario
    from mplambda_b200.planner import BatchedRRT, DeviceRRT
    for name in ("cubicles", "twistycool"):
        env_, robot_ = load_scene(name)
        sc_ = SE3_SCENES[name]
        g = M.SE3RigidBodyScenario(env_, robot_, sc_["goal"], sc_["min"], sc_["max"], 0.1, device=local_rank)
        tg, tc, td, same = [], [], [], True
        for seed in (1, 2, 3):
            rg = BatchedRRT(g, M.Nigh(M.nn.SE3, device=local_rank), sc_["start"], batch=2048, seed=seed)
            pg = rg.solve(time_limit=30)
            # (the two GPU planners back to back: after the CPU planner's half second the GPU has idled down, and a
            # 7 ms run measured right behind it is mostly the clocks ramping up)
            rd = DeviceRRT(g, M.Nigh(M.nn.SE3, device=local_rank), sc_["start"], batch=2048, seed=seed)
            pd = rd.solve(time_limit=30)
            rc = BatchedRRT(CpuScenario(env_, robot_, sc_, thr), CpuNN(thr), sc_["start"], batch=2048, seed=seed)
            pc = rc.solve(time_limit=120)
            tg.append(rg.stats["seconds"]); tc.append(rc.stats["seconds"])
            if pd is not None:
                td.append(rd.stats["seconds"])
            same = same and pg is not None and pc is not None and rg.n == rc.n and np.array_equal(pg, pc)
            rd.close()
        out["rrt_" + name] = {"metric": "RRT time-to-solution (batch-issuing PRRT loop, batch 2048, median of 3 seeds)",
                              "value": float(np.median(tg)), "unit": "s", "higher_is_better": False,
                              "device_resident_loop_s": float(np.median(td)) if td else None,
                              "cpu_baseline": {"value": float(np.median(tc)), "unit": "s", "cores": thr, "kind": "port",
                                               "sample": "the same batch-issuing planner (this repo's BatchedRRT, not the "
                                                         "reference's PRRT), seeds and batches, on the oracle"},
                              "identical_trees_and_paths": bool(same)}
        g.close()
    # ---- config 5 edges: 16,384 motions between random configurations (valid start), resolution 0.01
    A = Q[v][:1 << 14]
    B = fetch_configs(len(A), SEED + 1)
    f.check_edges(A[:256], B[:256])
    dt, ev = best(lambda: f.check_edges(A, B), reps=2)
    m = 256
    t = time.perf_counter(); w = O.FetchOracle(env, frame, 0.01).check_edges(A[:m], B[:m], threads=thr); dc = time.perf_counter() - t
    out["fetch_edges"] = {"metric": "Fetch edge validity checks/sec (resolution 0.01)", "value": len(A) / dt, "unit": "edges/s",
                          "edges": int(len(A)), "valid_fraction": float(ev.mean()), "ms": dt * 1e3,
                          "cpu_baseline": {"value": m / dc, "unit": "edges/s", "cores": thr, "kind": "port",
                                           "sample": "first %d edges" % m},
                          "verdict_mismatches_vs_cpu": int((w != ev[:m]).sum())}
    return out


def multi_gpu_extras(rank, world, local_rank):
    """BASELINE configs 3-5 at N > 1: fixed batches split over the ranks (strong scaling; the scene is replicated, a
    rank checks its contiguous share through the C ABI, no collective on the check path), and independent planner
    shards that exchange their solution over NCCL.  Times are max over ranks between two barriers."""
    import torch
    import torch.distributed as dist
    import mplambda_b200 as M
    from mplambda_b200 import sharding
    from mplambda_b200.planner import BatchedRRT, solve_sharded
    from mplambda_b200.workload import FETCH_TASKS, fetch_configs, frame_from_option
    out = {}

    def timed(fn, reps=3):
        ts = []
        for _ in range(reps):
            dist.barrier(); torch.cuda.synchronize()
            t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ts.append(float(tt.item()))
        return min(ts), r

    def total(x):
        tt = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt)
        return float(tt.item())

    # ---- config 3: 65,536-edge batches, cut where the summed step counts balance
    for name in ("cubicles", "apartment"):
        env, robot = load_scene(name)
        sc = SE3_SCENES[name]
        s = M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], 0.1, device=local_rank)
        P = se3_poses(8 << 16, sc["min"], sc["max"], SEED)
        A = P[s.check_states(P)][:1 << 16]
        B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
        lo, hi = sharding.balanced_edge_bounds(s.edge_steps(A, B), world)[rank]
        a, b = np.ascontiguousarray(A[lo:hi]), np.ascontiguousarray(B[lo:hi])
        s.check_edges(a, b)
        dt, v = timed(lambda: s.check_edges(a, b))
        out["edges_" + name] = {"metric": "edge validity checks/sec (isValid(from,to), resolution 0.1)", "unit": "edges/s",
                                "value": len(A) / dt, "edges": int(len(A)), "ms": dt * 1e3, "scaling": "strong",
                                "valid_fraction": total(v.sum()) / len(A),
                                "sharding": "contiguous shares balanced by step count (balanced_edge_bounds)"}
        s.close()
    # ---- config 4: Home, k-NN of the new nodes in the replicated tree + the k motions per node; new nodes split by rank
    sc = SE3_SCENES["home"]
    env, robot = load_scene("home")
    s = M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], 0.1, device=local_rank)
    keys = se3_poses(1 << 14, sc["min"], sc["max"], SEED)
    qs = se3_poses(1 << 16, sc["min"], sc["max"], SEED + 2)
    tree = keys[s.check_states(keys)]
    allnew = qs[s.check_states(qs)][:32768]
    lo, hi = sharding.shard_bounds(len(allnew), rank, world)
    new = np.ascontiguousarray(allnew[lo:hi])
    k = int(math.ceil(math.e * (7.0 / 6.0) * math.log(len(tree) + 1)))
    nn = M.Nigh(M.nn.SE3, device=local_rank)
    nn.insert(tree)
    cs = torch.cuda.current_stream()
    h_new = torch.from_numpy(new).pin_memory()
    d_new = torch.empty((len(new), 7), dtype=torch.float64, device="cuda")
    d_i = torch.empty((len(new), k), dtype=torch.int64, device="cuda")
    d_d = torch.empty((len(new), k), dtype=torch.float64, device="cuda")
    d_v = torch.empty(len(new) * k, dtype=torch.uint8, device="cuda")
    h_i = torch.empty((len(new), k), dtype=torch.int64).pin_memory()
    h_v = torch.empty(len(new) * k, dtype=torch.uint8).pin_memory()

    def near_and_connect():
        d_new.copy_(h_new, non_blocking=True)
        nn.knearest_device(d_new.data_ptr(), len(new), k, d_i.data_ptr(), d_d.data_ptr(), cs.cuda_stream)
        s.check_edges_indexed_device(d_new.data_ptr(), len(new), None, nn.keys_device(), nn.size(), d_i.data_ptr(),
                                     len(new) * k, d_v.data_ptr(), from_div=k, stream=cs.cuda_stream)
        h_i.copy_(d_i, non_blocking=True)
        h_v.copy_(d_v, non_blocking=True)
        cs.synchronize()
        return h_v.numpy().astype(bool)
    near_and_connect()
    dt, ev = timed(near_and_connect)
    out["knn_connect_home"] = {"metric": "new nodes/sec through k-NN + k motion checks (PCForest addNodeNear pattern)",
                               "unit": "nodes/s", "value": len(allnew) / dt, "new_nodes": int(len(allnew)), "k": k,
                               "tree_nodes": int(len(tree)), "ms": dt * 1e3, "scaling": "strong",
                               "edge_valid_fraction": total(ev.sum()) / (len(allnew) * k)}
    s.close()
    # ---- config 5: Fetch, 2^20 configurations split by rank
    env = np.load(os.path.join(ROOT, "tests", "golden", "scenes", "autolab.npz"))["env"]
    frame = frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"])
    f = M.FetchScenario(frame, env, checkResolution=0.01, device=local_rank)
    nq = 1 << 20
    lo, hi = sharding.shard_bounds(nq, rank, world)
    Q = fetch_configs(hi - lo, SEED, start=lo)
    f.check_states(Q)
    dt, v = timed(lambda: f.check_states(Q))
    out["fetch_states"] = {"metric": "Fetch validity checks/sec (FK + 24 MPR pairs + 12 shapes vs mesh)", "unit": "checks/s",
                           "value": nq / dt, "configs": nq, "ms": dt * 1e3, "scaling": "strong",
                           "valid_fraction": total(v.sum()) / nq}
    f.close()
    # ---- planner shards (the reference's lambdas, mpl_coordinator.cpp:591-600): one RRT per rank, seeded per rank, the
    #      Done flag and the winning path travel over NCCL (SolutionExchange) instead of the coordinator's TCP relays
    name = "twistycool"
    env_, robot_ = load_scene(name)
    sc_ = SE3_SCENES[name]
    g = M.SE3RigidBodyScenario(env_, robot_, sc_["goal"], sc_["min"], sc_["max"], 0.1, device=local_rank)
    ex = sharding.SolutionExchange(state_dim=7)
    spent = {"done_polls": 0, "done_s": 0.0, "publish_s": 0.0}
    any_done, publish = ex.any_done, ex.publish

    def any_done_timed(flag):
        t = time.perf_counter(); r = any_done(flag); spent["done_s"] += time.perf_counter() - t; spent["done_polls"] += 1
        return r

    def publish_timed(cost, path):
        t = time.perf_counter(); r = publish(cost, path); spent["publish_s"] += time.perf_counter() - t
        return r
    ex.any_done, ex.publish = any_done_timed, publish_timed
    ex.any_done(False)   # first collective of the group outside the timing
    spent.update(done_polls=0, done_s=0.0)
    secs = []
    for seed in (1, 2, 3):
        dist.barrier()
        t = time.perf_counter()
        cost, who, path, rrt = solve_sharded(
            lambda r: BatchedRRT(g, M.Nigh(M.nn.SE3, device=local_rank), sc_["start"], batch=2048, seed=1000 * seed + r),
            ex, time_limit=30, sync_every=1)
        secs.append(time.perf_counter() - t)
    ok = path is not None and bool(g.check_edges(path[:-1], path[1:]).all()) and bool(g.isGoal(path[-1]))
    out["sharded_rrt_" + name] = {"metric": "time to first solution, one RRT shard per GPU (seeded per rank), Done flag and "
                                            "winning path exchanged over NCCL", "unit": "s", "higher_is_better": False,
                                  "value": float(np.median(secs)), "shards": world, "last_winner_rank": int(who),
                                  "last_path_waypoints": 0 if path is None else int(len(path)), "last_path_valid": ok,
                                  "exchange_us_per_done_poll": 1e6 * spent["done_s"] / max(1, spent["done_polls"]),
                                  "exchange_us_per_publish": 1e6 * spent["publish_s"] / 3,
                                  "note": "all_reduce(MAX) of one int32 after every planner iteration; all_gather of (cost, "
                                          "length) + broadcast of the winner's waypoints once per solve"}
    g.close()
    return out


def contract_config(workload, env, robot):
    return {"workload": "%s SE(3) rigid body, 2^20 uniformly sampled poses per GPU per step "
                        "(randomize.hpp distribution, Philox seed 0x6d706c62)" % workload,
            "env_tris": int(len(env)), "robot_tris": int(len(robot)), "poses_per_step_per_gpu": N_POSES,
            "l2": "inputs rotate over %d distinct 58.7 MB pose batches (> 126 MB L2); the BVH itself is "
                  "L2/L1 resident by design" % ROTATE,
            "sharding": "pose stream split by rank, BVH replicated, no collective"}


def run_reference(args, rank, world):
    """The CPU path on the contract workload: the same 2^20-pose batches, one batch per step, all host cores."""
    if rank != 0:
        return
    from oracle import oracle_py as O
    env, robot = load_scene(args.workload)
    sc = SE3_SCENES[args.workload]
    orc = O.SE3Oracle(env, robot)
    threads = host_threads()
    sample = args.ref_sample
    batches = [se3_poses(sample, sc["min"], sc["max"], SEED, start=r * N_POSES) for r in range(ROTATE)]
    for w in range(args.warmup):
        orc.check_states(batches[w % ROTATE][:min(sample, 65536)], O.BVH, threads=threads)
    t = time.perf_counter()
    for k in range(args.steps):
        orc.check_states(batches[k % ROTATE], O.BVH, threads=threads)
    dt = time.perf_counter() - t
    v = args.steps * sample / dt
    cfg = contract_config(args.workload, env, robot)
    if sample != N_POSES:
        cfg["workload"] += " [--ref-sample: %d-pose sample of each step]" % sample
    print(json.dumps({
        "impl": "reference", "metric": "collision checks/sec (pose batch)", "value": v, "unit": "checks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": "checks/s", "cores": threads, "kind": "port",
                         "sample": "%d poses per step (the whole step), oracle O2 (FP64 restatement of FCL's OBB traversal + "
                                   "17-axis tri-tri; FCL itself is not installable here); rank 0's host cores only, whatever N"
                                   % sample},
        "e2e": {"value": v, "unit": "checks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mplb", choices=["mplb", "reference"])
    ap.add_argument("--workload", default="alpha15", choices=sorted(SE3_SCENES))
    ap.add_argument("--ref-sample", type=int, default=N_POSES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="leave the pinned host batches where the OS puts them")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "mplb" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if world > 1:
        # torchrun exports OMP_NUM_THREADS=1; libmplb's host side (edge step planning, copies out of pageable memory)
        # uses OpenMP: give every rank its share of the host cores (set before libgomp is loaded)
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        os.environ["OMP_NUM_THREADS"] = str(max(1, host_threads() // max(1, local_world)))

    import torch
    import torch.distributed as dist
    import mplambda_b200 as M

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmplb has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    env, robot = load_scene(args.workload)
    sc = SE3_SCENES[args.workload]
    scen = M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], 0.1, device=local_rank)

    # rank r owns poses [r*ROTATE*N, (r+1)*ROTATE*N) of the seeded stream: ROTATE batches of 2^20
    base = rank * ROTATE * N_POSES
    host = [se3_poses(N_POSES, sc["min"], sc["max"], SEED, start=base + r * N_POSES) for r in range(ROTATE)]
    dev = [torch.from_numpy(h).cuda() for h in host]
    out = torch.empty(N_POSES, dtype=torch.uint8, device="cuda")
    with HostMemoryNearGpu(local_rank, not args.no_numa) as placement:
        pinned = [torch.from_numpy(h).pin_memory() for h in host]
        pinned_out = [torch.empty(N_POSES, dtype=torch.uint8).pin_memory() for _ in range(2)]
    if world > 1:
        print("rank %d: %s" % (rank, placement.note), file=sys.stderr)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Device-resident steps go out alternately on two streams (mplb.h: every mplb_check_states_device call in flight has
    # a work cursor of its own): the SMs one launch has drained start on the next batch while its last warps finish.
    # The single-stream figure is measured next to it and is what the roofline block's launch duration comes from.
    stream2 = torch.cuda.Stream()
    out2 = torch.empty_like(out)
    lanes = [(stream, out), (stream2, out2)]

    def step_device(k, two_streams=False):
        st_, out_ = lanes[k % 2 if two_streams else 0]
        scen.check_states_device(dev[k % ROTATE].data_ptr(), N_POSES, out_.data_ptr(), st_.cuda_stream)

    def step_e2e(k, slot=0):
        scen.check_states_ptr(pinned[k % ROTATE].data_ptr(), N_POSES, pinned_out[slot].data_ptr())

    def steps_e2e_two_callers(n):
        """n calls issued by two host threads (ctypes releases the GIL): step k goes to caller k % 2."""
        def caller(t):
            for k in range(t, n, 2):
                step_e2e(k, t)
        th = [threading.Thread(target=caller, args=(t,)) for t in range(2)]
        [t.start() for t in th]
        [t.join() for t in th]

    def step_sampled(k):
        # the poses of batch k % ROTATE drawn on the device (the same Philox stream as `host`), checked, verdicts back
        scen.sample_states_device(SEED, base + (k % ROTATE) * N_POSES, N_POSES, sampled.data_ptr(), stream.cuda_stream)
        scen.check_states_device(sampled.data_ptr(), N_POSES, out.data_ptr(), stream.cuda_stream)
        pinned_out[0].copy_(out, non_blocking=True)
        stream.synchronize()
    sampled = torch.empty((N_POSES, 7), dtype=torch.float64, device="cuda")

    # setup, not a step: every pinned batch crosses the link once so that none of them meets the DMA engine for the
    # first time inside the timed region (the first transfers from a fresh pinned allocation run at ~70 % speed)
    touch = torch.empty_like(dev[0])
    for r in range(ROTATE):
        touch.copy_(pinned[r], non_blocking=True)
    torch.cuda.synchronize()
    del touch
    for k in range(args.warmup):
        step_device(k)
        step_device(k, True)
        step_e2e(k)
        step_sampled(k)
    steps_e2e_two_callers(4)
    barrier()
    scen.stats(reset=True)

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- device-resident: K steps, CUDA events on the launching stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for k in range(args.steps):
        step_device(k)
        ev[k + 1].record(stream)
    barrier()
    serial_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    st = scen.stats(reset=True)
    # ---- the same K steps, alternately on two streams (the headline `value`)
    ev_a, ev_b, ev_join = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event())
    barrier()
    ev_a.record(stream)
    stream2.wait_event(ev_a)
    for k in range(args.steps):
        step_device(k, True)
    ev_join.record(stream2)
    stream.wait_event(ev_join)
    ev_b.record(stream)
    barrier()
    total_ms = ev_a.elapsed_time(ev_b)
    st_two = scen.stats(reset=True)
    # ---- end to end through the C ABI with host buffers: two callers, then one
    barrier()
    t0 = time.perf_counter()
    steps_e2e_two_callers(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e(k)
    barrier()
    e2e1_s = time.perf_counter() - t0
    st_e2e = scen.stats(reset=True)
    # ---- the poses never cross PCIe: sampled on the device, verdicts back
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_sampled(k)
    barrier()
    samp_s = time.perf_counter() - t0
    clocks = sampler.stop()
    same_as_host = bool((pinned_out[0].numpy() == scen.check_states(host[(args.steps - 1) % ROTATE]).astype(np.uint8)).all())
    # context for the e2e number: what the host link itself sustains for this step's input
    hev0, hev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch = torch.empty_like(dev[0])
    scratch.copy_(pinned[0], non_blocking=True)
    torch.cuda.synchronize()
    hev0.record(stream)
    for r in range(4):
        scratch.copy_(pinned[r % ROTATE], non_blocking=True)
    hev1.record(stream)
    torch.cuda.synchronize()
    h2d_ms = hev0.elapsed_time(hev1) / 4
    valid_frac = float(out.float().mean().item())

    t = torch.tensor([total_ms, e2e_s * 1e3, e2e1_s * 1e3, samp_s * 1e3, serial_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e1_ms, samp_ms, serial_ms = t.tolist()

    multi = None
    if world > 1 and not args.no_extras:
        try:
            multi = multi_gpu_extras(rank, world, local_rank)
        except Exception as e:  # the contract line must survive an extras failure
            multi = {"error": repr(e)}
            try:
                dist.barrier()
            except Exception:
                pass

    if rank == 0:
        units = args.steps * N_POSES * world
        value = units / (total_ms * 1e-3)
        line = {
            "metric": "collision checks/sec (pose batch)", "value": value, "unit": "checks/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
            "data": "synthetic",
            "config": contract_config(args.workload, env, robot),
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "checks/s", "h2d_bytes_per_step": N_POSES * 56,
                    "d2h_bytes_per_step": N_POSES, "ms_per_step": e2e_ms / args.steps,
                    "api": "mplb_check_states (pinned host buffers), calls issued by two host threads per GPU like the "
                           "reference's OpenMP planner threads (prrt.hpp:140-152); every call copies its 2^20 poses in "
                           "and its verdicts out",
                    "single_caller": {"value": units / (e2e1_ms * 1e-3), "ms_per_step": e2e1_ms / args.steps},
                    "host_memory": placement.note,
                    "h2d_copy_alone_ms": h2d_ms, "h2d_gbs": N_POSES * 56 / (h2d_ms * 1e-3) / 1e9,
                    "note": "the 56 B/pose input copy is the longer leg; the kernel runs underneath it and picks states up "
                            "as the copy's watermark advances"},
            "e2e_device_sampled": {"value": units / (samp_ms * 1e-3), "unit": "checks/s", "ms_per_step": samp_ms / args.steps,
                                   "h2d_bytes_per_step": 0, "d2h_bytes_per_step": N_POSES,
                                   "api": "mplb_sample_states_device + mplb_check_states_device + verdict read-back: the "
                                          "planner's samples are drawn on the device (same Philox stream), only verdicts "
                                          "cross PCIe", "verdicts_equal_host_batch": same_as_host},
            "issue": "the K device-resident steps go out alternately on two CUDA streams (each call has a work cursor of its "
                     "own): the SMs one launch has drained take the next batch's CTAs while its last warps finish",
            "single_stream": {"value": units / (serial_ms * 1e-3), "ms_per_step": serial_ms / args.steps,
                              "note": "the same K steps back to back on one stream; the roofline block's launch duration"},
            "gpu_launches": int(st_two["launches"]),
            "clocks": clocks,
            "valid_fraction": valid_frac,
            "device_counters_per_check": {"bv_tests": st["bv_tests"] / (args.steps * N_POSES),
                                          "tri_tests": st["tri_tests"] / (args.steps * N_POSES),
                                          "exact_fp64_tests": st["exact_tests"] / (args.steps * N_POSES)},
            "stream_stalls": int(st_e2e["stream_stalls"]),
        }
        nbv = st["bv_tests"] / (args.steps * N_POSES)
        ntri = st["tri_tests"] / (args.steps * N_POSES)
        # north_star: "tolerance-band disagreements counted and reported".  Verdicts are bit-exact by construction except
        # where the device's sin/acos (edge samples, <= 2 ulp from glibc) could flip an exact triangle-pair decision: one
        # untimed pass with margin tracking on reports how close any exact decision of this step's batch came to that.
        scen.track_margins(True)
        scen.stats(reset=True)
        scen.check_states_device(dev[0].data_ptr(), N_POSES, out.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        mst = scen.stats(reset=True)
        scen.track_margins(False)
        line["near_band"] = {"min_relative_margin_of_exact_decisions": mst["min_exact_margin"],
                             "exact_decisions": int(mst["exact_tests"]), "band": 1e-13,
                             "decisions_inside_band": 0 if mst["min_exact_margin"] > 1e-13 else None,
                             "note": "relative margin |G| of the FP64 triangle-pair decisions of one 2^20-pose batch (G: largest "
                                     "separating gap over |axis|_1 x largest coordinate; disjoint iff G > 0); pose checks use no "
                                     "transcendental at all, edge samples use the device's sin/acos (<= 2 ulp from glibc, ~1e-16 "
                                     "relative): a verdict can differ from the CPU path's only if a margin is inside the band"}
        if not args.no_cpu_baseline:
            cb, cpu_valid, ncpu = cpu_baseline(env, robot, host[0], 0, args.cpu_seconds)
            got = scen.check_states(host[0][:ncpu])
            cb["verdict_mismatches_vs_gpu"] = int((got != cpu_valid).sum())
            line["cpu_baseline"] = cb
        kms = float(np.mean(kernel_ms))
        rl = hbm_roofline("checkStatesKernel", 28 + 1 + 120 * nbv + 72 * ntri, N_POSES, kms,
                          "algorithmic bytes = what the device's own traversal reads (%.1f box tests x 120 B, %.2f triangle "
                          "tests x 72 B per check) if nothing were cached; the hierarchy is L1/L2 resident, DRAM sees the poses "
                          "in and the verdicts out (`traffic`), so the kernel is bound by the SM, not by HBM: see "
                          "binding_units" % (nbv, ntri), ncu_facts(args.workload))
        if "cpu_baseline" in line:
            cbv = line["cpu_baseline"]
            rl["frac_with_reference_order_counts"] = (29 + 120 * cbv["n_bv_per_check"] + 72 * cbv["n_tri_per_check"]) * N_POSES \
                / (kms * 1e-3) / 1e9 / rl["peak"]
        line["roofline"] = rl
        if world == 1 and not args.no_extras and not args.no_cpu_baseline:
            try:
                line["extras"] = extras(local_rank, args.cpu_seconds)
            except Exception as e:  # the contract line must survive an extras failure
                import traceback
                line["extras"] = {"error": repr(e), "where": traceback.format_exc()[-600:]}
        if multi is not None:
            line["multi_gpu"] = multi
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
