#!/usr/bin/env python
"""bench.py -- throughput of the experience data path on B200 (contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--only c2|c3|c4]

Headline workload (BASELINE.json configs[1], "C2"): PriorityReplayBuffer, capacity 1M (2^20 leaves),
40-byte rows (obs 4xf32, action i32, reward f32, next obs 4xf32), batch 256, alpha 0.6, beta 0.4.
One step = sample(256) followed by update_priorities(256 fresh priorities).  metric = samples/s.

  value     device-resident: priorities already in HBM, results stay in HBM; CUDA events per step
  e2e       the same step through the C ABI's *_host entry points with pinned HOST buffers
            (H2D of indices+priorities, D2H of indices, weights and rows inside the timed region)
  roofline  dominant kernel of the step against the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle (port of the reference's sequential algorithm) on one host core
  also      BASELINE configs[2] (Atari-shaped PER, gather-bound) and configs[3] (GAE 4096x2048),
            each with its own value / e2e / roofline / cpu_baseline (N=1 only)

L2 hygiene: before EVERY timed iteration a 256 MiB buffer is overwritten and a second 256 MiB buffer
is read (not timed; the read leaves the L2 full of clean lines, so the timed kernels do not pay for
the flush's write-backs); the step time is the sum of the per-iteration CUDA-event durations on the
launching stream.

--impl reference times the reference's CPU path for the same step: MindSpore is not installable
here, so this is the oracle port of its sequential CPU kernels (oracle/mrl_oracle.c).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALPHA, BETA = 0.6, 0.4
C2_ROWS = [16, 4, 4, 16]                 # R = 40
C3_ROWS = [28224, 4, 4, 1, 28224]        # R = 56457 (obs, action, reward, done, next obs)
L_1M = 20


def ncu_traffic(key):
    """per-launch DRAM bytes of a kernel from the committed ncu capture (profiles/r1_traffic.json)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        return t[key]["dram_traffic_bytes"]
    except Exception:
        return None


def ncu_latency_evidence(key):
    """what SURVEY 8(d) asks for the latency-bound kernels instead of a %-of-peak target: L2 hit rate and warp
    stall reasons of the committed ncu capture (profiles/r1_traffic.json; cold caches, serialised launches)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[key]
        return {k: t[k] for k in ("ncu_duration_us", "l2_hit_rate_pct", "warps_active_pct", "issue_active_pct",
                                  "stalls_per_issue", "source") if k in t}
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- clocks ------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None and self._t is None:
            self._stop.clear()
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- GPU helpers -------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # keep stdout to the one JSON line (NCCL prints its version banner there at VERSION/INFO)
            os.environ["NCCL_DEBUG"] = os.environ.get("MRL_NCCL_DEBUG", "WARN")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        from mindrl_b200 import _ffi
        self.ffi = _ffi
        self.L = _ffi.lib()
        _ffi.check(self.L.mrl_set_device(self.local_rank))
        self.stream = torch.cuda.current_stream().cuda_stream
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        self.flush_src = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
        self.flush_val = 0
        self.clocks = ClockSampler(self.local_rank)
        self.args = args
        self.align_buf = torch.zeros(1, dtype=torch.float32, device="cuda") if self.world > 1 else None

    def flush(self):
        # write 256 MiB (evicts everything), then read another 256 MiB so that the L2 is left holding
        # CLEAN lines: otherwise the timed kernel also pays for writing the flush's dirty lines back
        self.flush_val = (self.flush_val + 1) & 0xFF
        self.flush_buf.fill_(self.flush_val)
        self.flush_src.sum()
        if self.world > 1 and os.environ.get("MRL_BENCH_ALIGN", "0") == "1":
            # opt-in experiment: re-align the ranks after the flush with an (untimed) all-reduce.  Measured at
            # N=8: WORSE (33.8 us/step against 28.7 free-running) -- the ranks leave the all-reduce further
            # apart than the previous step's exchange left them.  Off by default.
            self.dist.all_reduce(self.align_buf)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, steps, warmup, sample_clocks=False):
        """-> (total ms over `steps` iterations, our kernel launches in the timed region)"""
        torch = self.torch
        for i in range(warmup):
            self.flush()
            step(i)
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        if sample_clocks:
            self.clocks.start()
        l0 = self.ffi.launch_count()
        for i in range(steps):
            self.flush()
            ev[i][0].record()
            step(warmup + i)
            ev[i][1].record()
        self.barrier()
        launches = self.ffi.launch_count() - l0
        if sample_clocks:
            self.clocks.stop()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        return self.max_over_ranks(total_ms), launches

    def timed_host(self, step, steps, warmup):
        """end-to-end: host wall clock around each iteration, stream drained inside"""
        for i in range(warmup):
            self.flush()
            step(i)
        self.barrier()
        total = 0.0
        for i in range(steps):
            self.flush()
            self.torch.cuda.synchronize()
            t0 = time.perf_counter()
            step(warmup + i)
            self.ffi.check(self.L.mrl_stream_sync(self.stream))
            total += time.perf_counter() - t0
        self.barrier()
        return self.max_over_ranks(total * 1e3)


def pinned(ctx, shape, dtype):
    t = ctx.torch.empty(shape, dtype=dtype).pin_memory()
    return t


def make_prb(ctx, capacity, rows, seed):
    """filled PER buffer: deterministic rows, priorities |N(0,1)| + 1e-3 through the update path"""
    torch, L, ffi = ctx.torch, ctx.L, ctx.ffi
    rb = np.asarray(rows, dtype=np.int64)
    h = C.c_int64(-1)
    ffi.check(L.mrl_prb_create(capacity, ALPHA, len(rows), rb.ctypes.data, seed, ctx.rank, C.byref(h)))
    cols = (C.c_void_p * len(rows))()
    ffi.check(L.mrl_prb_columns(h.value, cols))
    for c, r in enumerate(rows):
        ffi.check(L.mrl_fill_rows(cols[c], capacity, r, 1000 + 17 * ctx.rank + c, ctx.stream))
    # push the rows "onto themselves": advances the ring to full and sets every leaf to max_priority
    ffi.check(L.mrl_prb_push_batch(h.value, capacity, cols, ctx.stream))
    g = torch.Generator(device="cuda")
    g.manual_seed(1234 + ctx.rank)
    pr = torch.randn(capacity, device="cuda", generator=g).abs_() + 1e-3
    idx = torch.arange(capacity, device="cuda", dtype=torch.int64)
    for s in range(0, capacity, 8192):
        e = min(capacity, s + 8192)
        ffi.check(L.mrl_prb_update(h.value, idx[s:e].data_ptr(), pr[s:e].data_ptr(), e - s, ctx.stream))
    torch.cuda.synchronize()
    return h.value, cols


def bench_per(ctx, name, rows, batch, steps, warmup, primary):
    """PER sample + update_priorities loop (C2 / C3 / per-rank shard of C5)"""
    torch, L, ffi = ctx.torch, ctx.L, ctx.ffi
    cap = 1_000_000
    R = sum(rows)
    h, cols = make_prb(ctx, cap, rows, 7)
    st = ctx.stream
    idx = torch.empty(batch, dtype=torch.int64, device="cuda")
    w = torch.empty(batch, dtype=torch.float32, device="cuda")
    outs = [torch.empty((batch, r), dtype=torch.uint8, device="cuda") for r in rows]
    out_tab = ffi.ptr_table([o.data_ptr() for o in outs])
    pool = 64
    g = torch.Generator(device="cuda")
    g.manual_seed(99 + ctx.rank)
    prio = torch.randn((pool, batch), device="cuda", generator=g).abs_() + 1e-3
    prio_ptr = [prio[i].data_ptr() for i in range(pool)]
    sharded = ctx.world > 1
    stats = torch.empty(4, dtype=torch.float32, device="cuda")
    gstats = torch.empty((ctx.world, 4), dtype=torch.float32, device="cuda")
    p2p = sharded and os.environ.get("MRL_SHARD_EXCHANGE", "p2p") == "p2p" and ctx.world <= 32
    if p2p:  # peer mailboxes: the sample kernel exchanges the root statistics itself
        buf = C.create_string_buffer(64)
        ffi.check(L.mrl_prb_shard_init(h, ctx.rank, ctx.world, buf))
        parts = [None] * ctx.world
        ctx.dist.all_gather_object(parts, buf.raw)
        allh = b"".join(parts)
        ffi.check(L.mrl_prb_shard_connect(h, C.create_string_buffer(allh, len(allh))))
        ctx.dist.barrier()

    def step(i):
        if p2p:
            ffi.check(L.mrl_prb_sample_sharded_p2p(h, batch, BETA, None, idx.data_ptr(), w.data_ptr(),
                                                   out_tab, st))
        elif sharded:  # root stats all-gather (16 B/rank over NVLink), then local strata
            ffi.check(L.mrl_prb_root_stats(h, stats.data_ptr(), st))
            ctx.dist.all_gather_into_tensor(gstats, stats)
            ffi.check(L.mrl_prb_sample_sharded(h, batch, BETA, None, gstats.data_ptr(), ctx.world,
                                               idx.data_ptr(), w.data_ptr(), out_tab, st))
        else:
            ffi.check(L.mrl_prb_sample(h, batch, BETA, None, idx.data_ptr(), w.data_ptr(), out_tab, st))
        ffi.check(L.mrl_prb_update(h, idx.data_ptr(), prio_ptr[i % pool], batch, st))

    total_ms, launches = ctx.timed(step, steps, warmup, sample_clocks=primary)
    value = ctx.world * batch * steps / (total_ms * 1e-3)

    # ---- per-kernel durations for the roofline (same flush discipline) -------------------
    ksteps = max(20, min(steps, 300))
    hbm, peak_src = peaks()

    def only_sample(i):
        ffi.check(L.mrl_prb_sample(h, batch, BETA, None, idx.data_ptr(), w.data_ptr(), None, st))

    def only_update(i):
        ffi.check(L.mrl_prb_update(h, idx.data_ptr(), prio_ptr[i % pool], batch, st))

    t_sample, _ = ctx.timed(only_sample, ksteps, 3)
    t_update, _ = ctx.timed(only_update, ksteps, 3)
    kern = {
        "prb_sample_kernel(descent+weights)": {"ms": t_sample / ksteps, "bytes": (4 * L_1M + 16) * batch},
        "prb_update_kernel": {"ms": t_update / ksteps, "bytes": (24 * L_1M + 20) * batch},
    }
    if R > 512:  # wide columns are gathered by their own kernel: time it through a uniform-buffer view
        wide = [c for c, r in enumerate(rows) if r > 512]
        uh = C.c_int64(-1)
        rb = np.asarray([rows[c] for c in wide], dtype=np.int64)
        wcols = ffi.ptr_table([cols[c] for c in wide])
        ffi.check(L.mrl_urb_create(cap, len(wide), rb.ctypes.data, wcols, 0, C.byref(uh)))
        wide_tab = ffi.ptr_table([outs[c].data_ptr() for c in wide])

        def only_gather(i):
            ffi.check(L.mrl_urb_gather(uh.value, idx.data_ptr(), batch, wide_tab, st))

        t_gather, _ = ctx.timed(only_gather, ksteps, 3)
        kern["gather_bulk_kernel"] = {"ms": t_gather / ksteps, "bytes": 2 * sum(rows[c] for c in wide) * batch}
        ffi.check(L.mrl_urb_destroy(uh.value))
    else:
        kern["prb_sample_kernel(descent+weights)"]["bytes"] += 0
    # the roofline is quoted for the bandwidth-bound kernel of the step when there is one (the row gather
    # of the Atari-shaped buffer), else for the slowest (latency-bound) tree kernel
    dom = "gather_bulk_kernel" if "gather_bulk_kernel" in kern else max(kern, key=lambda k: kern[k]["ms"])
    ach = kern[dom]["bytes"] / (kern[dom]["ms"] * 1e-3) / 1e9
    tkey = ("c3:" if R > 512 else "c2:") + {"prb_sample_kernel(descent+weights)": "prb_sample_kernel",
                                             "prb_update_kernel": "prb_update_sub_kernel",
                                             "gather_bulk_kernel": "gather_bulk_kernel<4>"}.get(dom, dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s",
                "frac": ach / hbm, "traffic": ncu_traffic(tkey), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kern[dom]["bytes"],
                "kernel_ms": {k: v["ms"] for k, v in kern.items()},
                "ncu": {k: ncu_latency_evidence(("c3:" if R > 512 else "c2:") + k)
                        for k in ("prb_sample_kernel", "prb_update_sub_kernel")},
                "note": ("tree kernels are dependent-load latency bound (L2/HBM round trips), "
                         "not bandwidth bound" if R <= 512 else "row gather is HBM-bandwidth bound")}

    # ---- e2e through the *_host entry points with pinned host buffers --------------------
    h_idx = pinned(ctx, (batch,), torch.int64)
    h_w = pinned(ctx, (batch,), torch.float32)
    h_out = [pinned(ctx, (batch, r), torch.uint8) for r in rows]
    h_tab = ffi.ptr_table([o.data_ptr() for o in h_out])
    h_prio = pinned(ctx, (pool, batch), torch.float32)
    h_prio.copy_(prio.cpu())
    h_prio_ptr = [h_prio[i].data_ptr() for i in range(pool)]

    def step_host(i):
        if sharded and not p2p:
            ffi.check(L.mrl_prb_root_stats(h, stats.data_ptr(), st))
            ctx.dist.all_gather_into_tensor(gstats, stats)
        if p2p:
            ffi.check(L.mrl_prb_sample_sharded_p2p_host(h, batch, BETA, None, h_idx.data_ptr(),
                                                        h_w.data_ptr(), h_tab, st))
        else:
            ffi.check(L.mrl_prb_sample_host(h, batch, BETA, None, h_idx.data_ptr(), h_w.data_ptr(), h_tab, st))
        ffi.check(L.mrl_prb_update_host(h, h_idx.data_ptr(), h_prio_ptr[i % pool], batch, st))

    esteps = max(10, min(steps, 200 if R > 512 else 1000))
    e_ms = ctx.timed_host(step_host, esteps, 3)
    e2e = {"value": ctx.world * batch * esteps / (e_ms * 1e-3), "unit": "samples/s",
           "h2d_bytes_per_step": batch * 12, "d2h_bytes_per_step": batch * (12 + R),
           "steps": esteps, "ms_per_step": e_ms / esteps}
    ffi.check(L.mrl_prb_destroy(h))
    return {"workload": name, "metric": "per_sample_update_samples_per_s", "value": value,
            "unit": "samples/s", "ms_per_step": total_ms / steps, "steps": steps,
            "gpu_launches": launches, "roofline": roofline, "e2e": e2e,
            "algorithmic_bytes_per_sample": 2 * R + 28 * L_1M + 36}


def bench_urb(ctx, steps, warmup):
    """C1: the DQN loop's buffer work -- uniform buffer cap 100k, 40-byte rows, per step insert 1 row + sample 64"""
    torch, L, ffi = ctx.torch, ctx.L, ctx.ffi
    cap, batch, rows = 100_000, 64, C2_ROWS
    R = sum(rows)
    st = ctx.stream
    rb = np.asarray(rows, dtype=np.int64)
    h = C.c_int64(-1)
    ffi.check(L.mrl_urb_create(cap, len(rows), rb.ctypes.data, None, 5, C.byref(h)))
    cols = (C.c_void_p * len(rows))()
    ffi.check(L.mrl_urb_columns(h.value, cols))
    fill = [torch.randint(0, 256, (cap, r), dtype=torch.uint8, device="cuda") for r in rows]
    ffi.check(L.mrl_urb_append(h.value, cap, ffi.ptr_table([f.data_ptr() for f in fill]), st))
    pool = 64
    new_rows = [torch.randint(0, 256, (pool, r), dtype=torch.uint8, device="cuda") for r in rows]
    row_tabs = [ffi.ptr_table([c[i].data_ptr() for c in new_rows]) for i in range(pool)]
    slots = torch.empty(batch, dtype=torch.int64, device="cuda")
    outs = [torch.empty((batch, r), dtype=torch.uint8, device="cuda") for r in rows]
    out_tab = ffi.ptr_table([o.data_ptr() for o in outs])

    def step(i):
        ffi.check(L.mrl_urb_append(h.value, 1, row_tabs[i % pool], st))
        ffi.check(L.mrl_urb_sample(h.value, batch, slots.data_ptr(), out_tab, st))

    total_ms, launches = ctx.timed(step, steps, warmup)
    per = total_ms / steps
    hbm, peak_src = peaks()
    alg = 2 * R + (2 * R + 8) * batch
    ach = alg / (per * 1e-3) / 1e9
    h_rows = [pinned(ctx, (pool, r), torch.uint8) for r in rows]
    for a, b in zip(h_rows, new_rows):
        a.copy_(b.cpu())
    h_row_tabs = [ffi.ptr_table([c[i].data_ptr() for c in h_rows]) for i in range(pool)]
    h_slots = pinned(ctx, (batch,), torch.int64)
    h_out = [pinned(ctx, (batch, r), torch.uint8) for r in rows]
    h_tab = ffi.ptr_table([o.data_ptr() for o in h_out])

    def step_host(i):
        ffi.check(L.mrl_urb_append_host(h.value, 1, h_row_tabs[i % pool], st))
        ffi.check(L.mrl_urb_sample_host(h.value, batch, h_slots.data_ptr(), h_tab, st))

    esteps = max(10, min(steps, 1000))
    e_ms = ctx.timed_host(step_host, esteps, 3)
    ffi.check(L.mrl_urb_destroy(h.value))
    return {"workload": "C1 DQN-shaped uniform buffer cap 100k, 40-byte rows: insert 1 + sample 64 per step",
            "metric": "uniform_insert_sample_samples_per_s", "value": ctx.world * batch * steps / (total_ms * 1e-3),
            "unit": "samples/s", "ms_per_step": per, "steps": steps, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "copy_seg_kernel + urb_sample_fused_kernel", "achieved": ach,
                         "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg,
                         "note": "two launches of a few KB each: launch-latency bound, not bandwidth bound"},
            "e2e": {"value": ctx.world * batch * esteps / (e_ms * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": R, "d2h_bytes_per_step": batch * (8 + R), "steps": esteps,
                    "ms_per_step": e_ms / esteps}}


def bench_insert(ctx, steps, warmup):
    """bulk insert (BufferAppend / push with a batch dimension): 2048 Atari-shaped rows per step"""
    torch, L, ffi = ctx.torch, ctx.L, ctx.ffi
    cap, n, rows = 65536, 2048, C3_ROWS
    R = sum(rows)
    st = ctx.stream
    rb = np.asarray(rows, dtype=np.int64)
    src = [torch.randint(0, 256, (n, r), dtype=torch.uint8, device="cuda") for r in rows]
    src_tab = ffi.ptr_table([t.data_ptr() for t in src])
    hbm, peak_src = peaks()
    out = {}
    h = C.c_int64(-1)
    ffi.check(L.mrl_urb_create(cap, len(rows), rb.ctypes.data, None, 5, C.byref(h)))
    ms, launches = ctx.timed(lambda i: ffi.check(L.mrl_urb_append(h.value, n, src_tab, st)), steps, warmup)
    out["uniform_append"] = {"ms": ms / steps, "GBps": 2 * R * n / (ms / steps * 1e-3) / 1e9, "launches": launches}
    # e2e: the same rows from pinned host arrays (one staged copy + the ring scatter)
    h_src = [pinned(ctx, (n, r), torch.uint8) for r in rows]
    for a, b in zip(h_src, src):
        a.copy_(b.cpu())
    h_tab = ffi.ptr_table([t.data_ptr() for t in h_src])
    esteps = max(5, min(steps, 20))
    e_ms = ctx.timed_host(lambda i: ffi.check(L.mrl_urb_append_host(h.value, n, h_tab, st)), esteps, 3)
    e2e = {"value": ctx.world * n * esteps / (e_ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": R * n,
           "d2h_bytes_per_step": 0, "steps": esteps, "ms_per_step": e_ms / esteps}
    ffi.check(L.mrl_urb_destroy(h.value))
    ffi.check(L.mrl_prb_create(cap, ALPHA, len(rows), rb.ctypes.data, 7, 0, C.byref(h)))
    ms, launches = ctx.timed(lambda i: ffi.check(L.mrl_prb_push_batch(h.value, n, src_tab, st)), steps, warmup)
    out["priority_push_batch(rows + leaves + tree rebuild)"] = {
        "ms": ms / steps, "GBps": 2 * R * n / (ms / steps * 1e-3) / 1e9, "launches": launches}
    ffi.check(L.mrl_prb_destroy(h.value))
    for v in out.values():
        v["frac"] = v["GBps"] / hbm
    main = out["uniform_append"]
    return {"workload": f"insert: {n} Atari-shaped rows ({R} bytes each) per step into a {cap}-row ring",
            "metric": "insert_rows_per_s", "value": ctx.world * n / (main["ms"] * 1e-3), "unit": "rows/s",
            "ms_per_step": main["ms"], "steps": steps, "gpu_launches": main["launches"],
            "roofline": {"bound": "hbm", "kernel": "copy_seg_kernel", "achieved": main["GBps"], "peak": hbm,
                         "unit": "GB/s", "frac": main["frac"], "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": 2 * R * n},
            "e2e": e2e, "variants": out}


def cpu_urb(seconds=2.0):
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    cap, batch = 100_000, 64
    shapes, types = [(r,) for r in C2_ROWS], [np.uint8] * len(C2_ROWS)
    b = O.UniformBufferOracle(batch, cap, shapes, types, seed=5)
    b.insert([rng.integers(0, 256, size=(cap, r), dtype=np.uint8) for r in C2_ROWS])
    rows = [[rng.integers(0, 256, size=(r,), dtype=np.uint8) for r in C2_ROWS] for _ in range(64)]
    for i in range(10):
        b.insert(rows[i]); b.sample()
    t0 = time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < seconds:
        b.insert(rows[steps % 64]); b.sample()
        steps += 1
    dt = time.perf_counter() - t0
    return batch * steps / dt, steps, dt


def bench_gae(ctx, steps, warmup):
    """C4: 4096 envs x 2048 steps fp32, gamma .99 lambda .95, random done flags"""
    torch, L, ffi = ctx.torch, ctx.L, ctx.ffi
    T, B = 2048, 4096
    st = ctx.stream
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    r = torch.randn((T, B), device="cuda", generator=g)
    v = torch.randn((T + 1, B), device="cuda", generator=g)
    d = (torch.rand((T, B), device="cuda", generator=g) < 1 / 200).to(torch.uint8)
    adv = torch.empty((T, B), device="cuda")
    ret = torch.empty((T, B), device="cuda")
    hbm, peak_src = peaks()
    n = T * B
    out = {}

    def run(name, fn, bytes_per_elem, extra_bytes=0):
        ms, launches = ctx.timed(fn, steps, warmup)
        per = ms / steps
        ach = (bytes_per_elem * n + extra_bytes) / (per * 1e-3) / 1e9
        out[name] = {"elems_per_s": n / (per * 1e-3), "ms": per, "GBps": ach, "frac": ach / hbm,
                     "launches": launches}

    def gae_tm(i):
        ffi.check(L.mrl_gae(r.data_ptr(), v.data_ptr(), None, v[T].data_ptr(), d.data_ptr(),
                            adv.data_ptr(), ret.data_ptr(), T, B, 0.99, 0.95, 0, 0, st))

    def gae_tm_seq(i):
        ffi.check(L.mrl_gae(r.data_ptr(), v.data_ptr(), None, v[T].data_ptr(), d.data_ptr(),
                            adv.data_ptr(), ret.data_ptr(), T, B, 0.99, 0.95, 0, 1, st))

    def gae_bm(i):  # the same buffers read as [B envs, T steps] (PPO layout)
        ffi.check(L.mrl_gae(r.data_ptr(), v.data_ptr(), None, v[T].data_ptr(), d.data_ptr(),
                            adv.data_ptr(), ret.data_ptr(), T, B, 0.99, 0.95, 1, 0, st))

    def dr_tm(i):
        ffi.check(L.mrl_discounted_return(r.data_ptr(), d.data_ptr(), v[T].data_ptr(), adv.data_ptr(),
                                          T, B, 1, 0.99, 0, 0, st))

    def dr_bm(i):
        ffi.check(L.mrl_discounted_return(r.data_ptr(), d.data_ptr(), v[T].data_ptr(), adv.data_ptr(),
                                          T, B, 1, 0.99, 1, 0, st))

    # sustained: five buffer sets rotated (710 MB of GAE inputs+outputs, nothing reused out of the 126 MB L2),
    # launches back to back inside ONE event pair: no per-launch event overhead, but every launch also
    # carries the write-back of its predecessor's outputs
    sets = [(r, v, d, adv, ret)]
    for _ in range(4):
        sets.append((torch.randn((T, B), device="cuda", generator=g), torch.randn((T + 1, B), device="cuda", generator=g),
                     (torch.rand((T, B), device="cuda", generator=g) < 1 / 200).to(torch.uint8),
                     torch.empty((T, B), device="cuda"), torch.empty((T, B), device="cuda")))

    def sustained(name, call, bytes_per_elem):
        for s_ in sets:
            call(*s_)
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                for s_ in sets:
                    call(*s_)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (8 * len(sets))
            best = ms if best is None or ms < best else best
        out[name]["back_to_back_ms"] = best
        out[name]["back_to_back_GBps"] = (bytes_per_elem * n + 4 * B) / (best * 1e-3) / 1e9
        out[name]["back_to_back_frac"] = out[name]["back_to_back_GBps"] / hbm

    run("gae_time_major[T,B]", gae_tm, 17, 4 * B)
    run("gae_time_major_sequential(bit-exact)", gae_tm_seq, 17, 4 * B)
    run("gae_batch_major[B,T]", gae_bm, 17, 4 * B)
    run("discounted_return_time_major", dr_tm, 9, 4 * B)
    run("discounted_return_batch_major", dr_bm, 9, 4 * B)
    p_ = lambda x: x.data_ptr()  # noqa: E731
    sustained("gae_time_major[T,B]", lambda r_, v_, d_, a_, t_: ffi.check(L.mrl_gae(
        p_(r_), p_(v_), None, p_(v_[T]), p_(d_), p_(a_), p_(t_), T, B, 0.99, 0.95, 0, 0, st)), 17)
    sustained("gae_batch_major[B,T]", lambda r_, v_, d_, a_, t_: ffi.check(L.mrl_gae(
        p_(r_), p_(v_), None, p_(v_[T]), p_(d_), p_(a_), p_(t_), T, B, 0.99, 0.95, 1, 0, st)), 17)
    sustained("discounted_return_time_major", lambda r_, v_, d_, a_, t_: ffi.check(L.mrl_discounted_return(
        p_(r_), p_(d_), p_(v_[T]), p_(a_), T, B, 1, 0.99, 0, 0, st)), 9)
    sustained("discounted_return_batch_major", lambda r_, v_, d_, a_, t_: ffi.check(L.mrl_discounted_return(
        p_(r_), p_(d_), p_(v_[T]), p_(a_), T, B, 1, 0.99, 1, 0, st)), 9)
    del sets
    main = out["gae_time_major[T,B]"]
    roofline = {"bound": "hbm", "kernel": "scan_tm_strip_kernel<GAE, done, 32 cols, 16 steps/lane>", "achieved": main["GBps"],
                "peak": hbm, "unit": "GB/s", "frac": main["frac"],
                "traffic": ncu_traffic("c4:scan_tm_strip_kernel<1, 1, 32, 16, 1>"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": 17 * n + 4 * B}
    # e2e: host arrays in, host arrays out
    hr, hv = pinned(ctx, (T, B), torch.float32), pinned(ctx, (T + 1, B), torch.float32)
    hd = pinned(ctx, (T, B), torch.uint8)
    hadv, hret = pinned(ctx, (T, B), torch.float32), pinned(ctx, (T, B), torch.float32)
    hr.copy_(r.cpu()), hv.copy_(v.cpu()), hd.copy_(d.cpu())

    def gae_host(i):
        ffi.check(L.mrl_gae_host(hr.data_ptr(), hv.data_ptr(), None, hv[T].data_ptr(), hd.data_ptr(),
                                 hadv.data_ptr(), hret.data_ptr(), T, B, 0.99, 0.95, 0, 0, st))

    esteps = max(5, min(steps, 20))
    e_ms = ctx.timed_host(gae_host, esteps, 2)
    e2e = {"value": n * esteps / (e_ms * 1e-3), "unit": "elements/s", "h2d_bytes_per_step": 9 * n + 4 * B,
           "d2h_bytes_per_step": 8 * n, "steps": esteps, "ms_per_step": e_ms / esteps}
    bm = out["gae_batch_major[B,T]"]
    roofline["batch_major_kernel"] = {"kernel": "scan_bm_row_kernel<GAE, done>", "achieved": bm["GBps"], "frac": bm["frac"],
                                      "traffic": ncu_traffic("c4:scan_bm_row_kernel<1, 1, 0>")}
    roofline["back_to_back"] = {"achieved": main["back_to_back_GBps"], "frac": main["back_to_back_frac"],
                                "note": "5 rotated buffer sets, launches back to back in one event pair"}
    w = ctx.world
    e2e["value"] *= w
    name = "C4 GAE 4096 envs x 2048 steps fp32 gamma=0.99 lambda=0.95 done~Bernoulli(1/200)"
    if w > 1:
        name += f", env axis sharded over {w} ranks (4096 envs per rank, no communication; whole-job elements/s)"
    return {"workload": name,
            "metric": "gae_elements_per_s", "value": main["elems_per_s"] * w, "unit": "elements/s",
            "ms_per_step": main["ms"], "steps": steps, "roofline": roofline, "e2e": e2e,
            "variants": out}


# ---- CPU baselines (oracle port; the only place bench.py executes oracle/) ---------------
def cpu_per(rows, batch, capacity, seconds=8.0, max_steps=4000):
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    shapes = [(r,) for r in rows]
    types = [np.uint8] * len(rows)
    b = O.PriorityBufferOracle(ALPHA, capacity, batch, shapes, types, 7, 0)
    chunk = max(1, min(capacity, (64 << 20) // max(sum(rows), 1)))
    block = [rng.integers(0, 256, size=(chunk, r), dtype=np.uint8) for r in rows]
    for s in range(0, capacity, chunk):
        n = min(chunk, capacity - s)
        b.insert_batch([blk[:n] for blk in block])
    pr = (np.abs(rng.normal(size=capacity)) + 1e-3).astype(np.float32)
    b.update_priorities(np.arange(capacity), pr)
    prio = (np.abs(rng.normal(size=(64, batch))) + 1e-3).astype(np.float32)
    for i in range(3):
        out = b.sample(BETA)
        b.update_priorities(out[0], prio[i])
    t0 = time.perf_counter()
    steps = 0
    while steps < max_steps and time.perf_counter() - t0 < seconds:
        out = b.sample(BETA)
        b.update_priorities(out[0], prio[steps % 64])
        steps += 1
    dt = time.perf_counter() - t0
    return batch * steps / dt, steps, dt


def cpu_gae(T, B, reps=3):
    from oracle import oracle as O
    rng = np.random.default_rng(4)
    r = rng.normal(size=(T, B)).astype(np.float32)
    v = rng.normal(size=(T + 1, B)).astype(np.float32)
    d = rng.random((T, B)) < 1 / 200
    O.gae(r[:64], v[:64], None, v[64], d[:64], 0.99, 0.95)
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        O.gae(r, v[:T], None, v[T], d, 0.99, 0.95)
        best = min(best, time.perf_counter() - t0)
    t0 = time.perf_counter()
    O.discounted_return_numpy_loop(r, d, v[T], 0.99)
    t_np = time.perf_counter() - t0
    return T * B / best, T * B / t_np


def host_info():
    info = {"os_cpu_count": os.cpu_count(), "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS")}
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                info["cpu"] = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return info


def run_reference(args, emit):
    """the reference arm: the reference's sequential CPU algorithm (oracle port) on the host"""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    batch, cap = 256, 1_000_000
    rng = np.random.default_rng(0)
    shapes, types = [(r,) for r in C2_ROWS], [np.uint8] * 4
    b = O.PriorityBufferOracle(ALPHA, cap, batch, shapes, types, 7, 0)
    b.insert_batch([rng.integers(0, 256, size=(cap, r), dtype=np.uint8) for r in C2_ROWS])
    b.update_priorities(np.arange(cap), (np.abs(rng.normal(size=cap)) + 1e-3).astype(np.float32))
    prio = (np.abs(rng.normal(size=(64, batch))) + 1e-3).astype(np.float32)
    for i in range(max(args.warmup, 3)):
        out = b.sample(BETA)
        b.update_priorities(out[0], prio[i % 64])
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = b.sample(BETA)
        b.update_priorities(out[0], prio[i % 64])
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    line = {
        "impl": "reference", "metric": "per_sample_update_samples_per_s", "value": value,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 PER cap 1M (2^20 leaves) 40-byte rows batch 256 alpha 0.6 beta 0.4, "
                               "sample+update_priorities per step", "capacity": cap, "batch": batch},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": 1, "kind": "port",
                         "sample": f"{args.steps} steps of sample(256)+update(256) on the full 1M-leaf tree; "
                                   "MindSpore (the reference's kernels) is not installable here, so this is "
                                   "the oracle port of its sequential CPU algorithm",
                         "host": host_info()},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    # the driver reads ONE JSON line from stdout: native libraries (NCCL's version banner) write to
    # fd 1 too, so everything but that line is sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w")

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--only", default="all", choices=["all", "c1", "c2", "c3", "c4", "insert"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baselines")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args, emit)
        return

    ctx = Ctx(args)
    primary, also = None, []
    if args.only in ("all", "c2"):
        primary = bench_per(ctx, "C2 PER cap 1M (2^20 leaves) 40-byte rows batch 256 alpha 0.6 beta 0.4, "
                                 "sample+update_priorities per step", C2_ROWS, 256, args.steps, args.warmup, True)
    if ctx.world == 1 and args.only in ("all", "c3"):
        r = bench_per(ctx, "C3 Atari-shaped PER cap 1M obs 84x84x4 u8 (x2) batch 512", C3_ROWS, 512,
                      max(20, min(args.steps, 200)), args.warmup, primary is None)
        (also.append(r) if primary else None)
        primary = primary or r
    if ctx.world == 1 and args.only in ("all", "c4"):
        r = bench_gae(ctx, max(10, min(args.steps, 50)), args.warmup)
        (also.append(r) if primary else None)
        primary = primary or r
    if ctx.world == 1 and args.only in ("all", "c1"):
        r = bench_urb(ctx, max(20, min(args.steps, 500)), args.warmup)
        (also.append(r) if primary else None)
        primary = primary or r
    if ctx.world == 1 and args.only in ("all", "insert"):
        r = bench_insert(ctx, max(10, min(args.steps, 50)), args.warmup)
        (also.append(r) if primary else None)
        primary = primary or r
    if ctx.world > 1 and args.only == "all":
        r = bench_per(ctx, "C5 sharded PER 1M/rank batch 4096/rank, root-stat all-gather per sample",
                      C2_ROWS, 4096, max(20, min(args.steps, 300)), args.warmup, False)
        also.append(r)
        also.append(bench_gae(ctx, max(10, min(args.steps, 50)), args.warmup))

    if ctx.rank == 0:
        if not args.no_cpu:
            if "PER" in primary["workload"]:
                wide = "C3" in primary["workload"]
                v, n, dt = cpu_per(C3_ROWS if wide else C2_ROWS, 512 if wide else 256,
                                   32768 if wide else 1_000_000)
                primary["cpu_baseline"] = {
                    "value": v, "unit": "samples/s", "cores": 1, "kind": "port",
                    "sample": f"{n} steps in {dt:.1f}s of sample+update on the oracle, capacity "
                              f"{32768 if wide else 1_000_000}", "host": host_info()}
            elif primary["workload"].startswith("insert"):
                pass
            elif primary["workload"].startswith("C1"):
                v, n, dt = cpu_urb()
                primary["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": 1, "kind": "port",
                                           "sample": f"{n} steps of insert 1 + sample 64 in {dt:.1f}s on the oracle",
                                           "host": host_info()}
            else:
                g, npl = cpu_gae(2048, 4096)
                primary["cpu_baseline"] = {"value": g, "unit": "elements/s", "cores": 1, "kind": "port",
                                           "sample": "full 2048x4096 GAE, best of 3", "host": host_info()}
            for r in also:
                if r["workload"].startswith("C3"):
                    v, n, dt = cpu_per(C3_ROWS, 512, 32768, seconds=6.0)
                    r["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": 1, "kind": "port",
                                         "sample": f"{n} steps in {dt:.1f}s on a 32768-row oracle buffer (1.8 GB of rows)"}
                elif r["workload"].startswith("C1"):
                    v, n, dt = cpu_urb()
                    r["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": 1, "kind": "port",
                                         "sample": f"{n} steps of insert 1 + sample 64 in {dt:.1f}s on the oracle "
                                                   "(through its Python wrapper)"}
                elif r["workload"].startswith("C4"):
                    g, npl = cpu_gae(2048, 4096)
                    r["cpu_baseline"] = {"value": g, "unit": "elements/s", "cores": 1, "kind": "port",
                                         "sample": "full 2048x4096 GAE on the C oracle, best of 3",
                                         "numpy_loop_discounted_return_elements_per_s": npl}
        line = {
            "metric": primary["metric"], "value": primary["value"], "unit": primary["unit"],
            "n_gpus": ctx.world, "steps": primary["steps"], "warmup": args.warmup,
            "ms_per_step": primary["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": primary["workload"], "per_rank": True,
                       "l2": "flushed before every timed iteration (256 MiB written, then 256 MiB read; untimed); "
                             "step time = sum of per-iteration CUDA-event durations on the launching stream"
                             + ("; ranks re-aligned after every flush by an untimed 4-byte NCCL all-reduce"
                                if ctx.world > 1 and os.environ.get("MRL_BENCH_ALIGN", "0") == "1" else ""),
                       "sharding": ("per-rank sub-buffers (1M rows each); root statistics exchanged inside the sample "
                                    "kernel by peer stores over NVLink (MRL_SHARD_EXCHANGE=nccl: NCCL all-gather)"
                                    if ctx.world > 1 else "single GPU")},
            "clocks": ctx.clocks.summary(),
            "e2e": primary["e2e"], "gpu_launches": primary.get("gpu_launches"),
            "roofline": primary["roofline"], "cpu_baseline": primary.get("cpu_baseline"),
            "algorithmic_bytes_per_sample": primary.get("algorithmic_bytes_per_sample"),
            "variants": primary.get("variants"),
            "also": also,
        }
        emit(line)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
