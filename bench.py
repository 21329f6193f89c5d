#!/usr/bin/env python
"""bench.py -- throughput of the experience data path on B200 (contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--only c1|c2|c3|c4|c5|insert]

Headline workload (BASELINE.json configs[1], "C2"): PriorityReplayBuffer, capacity 1M (2^20 leaves),
40-byte rows (obs 4xf32, action i32, reward f32, next obs 4xf32), batch 256, alpha 0.6, beta 0.4.
One step = sample(256) followed by update_priorities(256 fresh priorities).  metric = samples/s.

  value     device-resident: priorities already in HBM, results stay in HBM; CUDA events per step
  e2e       the same step through the C ABI's *_host entry points with pinned HOST buffers
            (H2D of indices+priorities, D2H of indices, weights and rows inside the timed region)
  roofline  dominant kernel of the step against the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle (port of the reference's sequential algorithm) on one host core, the loop
            timed inside the C library
  parity_check  BEFORE the timed loops, every rank runs sample+update steps with injected uniforms on the very
            path that is timed next (sharded exchange included) against the CPU oracle: indices, weight bits,
            rows, whole tree.  A mismatch ends the run non-zero.
  exchange  (N > 1) the peer-mailbox exchange of the root statistics: protocol, exchange_wait_us per sample
  c5        BASELINE configs[4] at THIS N (N = 1: one shard at batch 4096), and at N > 1 the same step un-sharded
            in the same run with the resulting weak-scaling efficiency
  also      BASELINE configs[2] (Atari-shaped PER, gather-bound), configs[3] (GAE 4096x2048, with the
            element-wise error histogram of the parallel kernels), configs[0] (DQN-shaped uniform buffer)
            and bulk insert, each with its own value / e2e / roofline / cpu_baseline

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
    """per-launch DRAM bytes of a kernel from the committed ncu capture (profiles/r2_traffic.json)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        return t[key]["dram_traffic_bytes"]
    except Exception:
        return None


def ncu_latency_evidence(key):
    """what SURVEY 8(d) asks for the latency-bound kernels instead of a %-of-peak target: L2 hit rate and warp
    stall reasons of the committed ncu capture (profiles/r2_traffic.json; cold caches, serialised launches)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))[key]
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
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample_once(self):
        """one NVML reading of the SM clock and the clock-event reasons, taken by the CALLING thread.
        Called from inside the timed loop at iteration boundaries, in front of the (untimed) L2 flush: the host is
        several iterations ahead of the GPU there, so the GPU is under the loop's load while NVML reads it, and the
        pause falls between two timed regions, never inside one.  (A sampling THREAD, as in round 1, made the
        20-step runs at N = 8 up to 30 % slower: eight processes' NVML calls serialise in the driver and stall the
        launches of the process that shares them -- measured: 78.7 M samples/s with the thread against 110.3 M in the
        same run without it.)"""
        if self.nv is None:
            return
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

    def start(self):
        pass

    def stop(self):
        pass

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
            # (NCCL_DEBUG is left as the caller set it: main() already sends everything but the JSON line to stderr)
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
        self.align = None  # handle of the sharded buffer whose peers flush() waits for (see there)

    def flush(self):
        # write 256 MiB (evicts everything), then read another 256 MiB so that the L2 is left holding
        # CLEAN lines: otherwise the timed kernel also pays for writing the flush's dirty lines back
        self.flush_val = (self.flush_val + 1) & 0xFF
        self.flush_buf.fill_(self.flush_val)
        self.flush_src.sum()
        if self.world > 1 and self.align is not None:
            # Part of the (untimed) preparation of a timed step at N > 1: wait, through the library's own
            # mrl_prb_shard_wait, until the peers' publications the coming sample needs have arrived.  The sharded step
            # is lock-step across ranks with one step of slack, so without this every microsecond by which one rank's
            # UNTIMED flush runs longer than another's comes back as a wait inside the other ranks' TIMED sample
            # (measured at N = 8: up to 2.7 us per step on the rank that waits most, i.e. flush jitter billed to the
            # workload).  A learner has its own forward/backward pass in this place.  The loop without it is reported as
            # free_running.  (An NCCL all-reduce here instead costs the step that follows it 1.7 us: measured at N = 2.)
            self.ffi.check(self.L.mrl_prb_shard_wait(self.align, self.stream))

    def barrier(self):
        # (device work first: a rank must not pass the barrier -- and, say, free a mailbox its peers' courier
        # kernels still write into -- while another rank's kernels are in flight)
        self.torch.cuda.synchronize()
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
        l0 = self.ffi.launch_count()
        every = max(1, steps // 5)
        for i in range(steps):
            if sample_clocks and i % every == every // 2:
                self.clocks.sample_once()  # (see there: between two timed regions, GPU under load)
            self.flush()
            ev[i][0].record()
            step(warmup + i)
            ev[i][1].record()
        self.barrier()
        launches = self.ffi.launch_count() - l0
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
    return h.value, cols, pr


def build_parity_oracle(ctx, capacity, rows, init_prio, batch):
    """the CPU oracle holding what make_prb put into the GPU buffer: the same rows (the fill pattern is shared
    arithmetic, include/mrl_arith.h) when they are narrow, 8-byte stand-in rows carrying the slot id when they are
    Atari-sized (56 GB do not fit the host: those rows are checked through the slot id every row embeds)"""
    from oracle import oracle as O
    narrow = sum(rows) <= 512
    orows = rows if narrow else [8]
    o = O.PriorityBufferOracle(ALPHA, capacity, batch, [(r,) for r in orows], [np.uint8] * len(orows), 7, ctx.rank)
    if narrow:
        o.insert_batch([O.fill_rows(capacity, r, 1000 + 17 * ctx.rank + c) for c, r in enumerate(rows)])
    else:
        o.insert_batch([np.arange(capacity, dtype=np.int64).view(np.uint8).reshape(capacity, 8)])
    o.update_priorities(np.arange(capacity), init_prio.cpu().numpy())
    return o, narrow


def parity_check(ctx, h, o, narrow, rows, batch, sample_fn, steps=4):
    """BEFORE the timed loops: `steps` sample+update steps with injected uniforms on the GPU path this bench is
    about to time (sharded exchange included) against the CPU oracle -- indices, weights (bit patterns), rows and,
    at the end, the whole sum / min tree.  At N > 1 the oracle gets the global statistics from an all-gather of
    every rank's root statistics (NCCL), i.e. independently of the peer mailboxes under test."""
    torch, L, ffi = ctx.torch, ctx.L, ctx.ffi
    st = ctx.stream
    rng = np.random.default_rng(4242 + ctx.rank)
    idx = torch.empty(batch, dtype=torch.int64, device="cuda")
    w = torch.empty(batch, dtype=torch.float32, device="cuda")
    outs = [torch.empty((batch, r), dtype=torch.uint8, device="cuda") for r in rows]
    out_tab = ffi.ptr_table([x.data_ptr() for x in outs])
    stats = torch.empty(4, dtype=torch.float32, device="cuda")
    gstats = torch.empty((ctx.world, 4), dtype=torch.float32, device="cuda")
    ok, detail = True, []
    for k in range(steps):
        u = rng.random(batch).astype(np.float32)
        ud = torch.as_tensor(u).cuda()
        g = None
        if ctx.world > 1:
            ffi.check(L.mrl_prb_root_stats(h, stats.data_ptr(), st))
            ctx.dist.all_gather_into_tensor(gstats, stats)
            g = gstats.cpu().numpy()
        sample_fn(ud.data_ptr(), idx, w, out_tab)
        want = o.sample(BETA, uniforms=u, gstats=g)
        gi, gw = idx.cpu().numpy(), w.cpu().numpy()
        step_ok = np.array_equal(gi, want[0]) and np.array_equal(gw.view(np.uint32), want[1].view(np.uint32))
        if narrow:
            step_ok = step_ok and all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(outs, want[2:]))
        else:  # every row starts with its slot id (mrl_fill_rows)
            for a, r in zip(outs, rows):
                if r >= 8:
                    step_ok = step_ok and np.array_equal(a[:, :8].contiguous().cpu().numpy().view(np.int64).reshape(-1), gi)
        pr = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
        ffi.check(L.mrl_prb_update(h, idx.data_ptr(), torch.as_tensor(pr).cuda().data_ptr(), batch, st))
        torch.cuda.synchronize()
        o.update_priorities(want[0], pr)
        ok = ok and bool(step_ok)
        if not step_ok:
            detail.append(f"step {k}: {int((gi != want[0]).sum())} indices, "
                          f"{int((gw.view(np.uint32) != want[1].view(np.uint32)).sum())} weights differ")
    cap2 = 1 << L_1M
    gs, gm = np.empty(2 * cap2, np.float32), np.empty(2 * cap2, np.float32)
    ffi.check(L.mrl_prb_tree_dump(h, gs.ctypes.data, gm.ctypes.data, st))
    os_, om = o.tree()
    tree_ok = np.array_equal(gs[1:], os_[1:]) and np.array_equal(gm[1:], om[1:])
    ok = ok and bool(tree_ok)
    if ctx.world > 1:
        flag = torch.tensor([1 if ok else 0], device="cuda")
        ctx.dist.all_reduce(flag, op=ctx.dist.ReduceOp.MIN)
        ok = bool(flag.item() == 1)
    return {"world": ctx.world, "steps": steps, "batch": batch, "ok": ok, "tree_bit_exact": bool(tree_ok),
            "checked": "indices, weight bits, " + ("rows" if narrow else "slot id embedded in every wide row")
                       + ", sum/min tree after the last update; oracle = CPU restatement, global statistics from "
                         "an NCCL all-gather" + ("" if not detail else "; " + "; ".join(detail))}


def bench_per(ctx, name, rows, batch, steps, warmup, primary, sharded=None, parity=True, details=True):
    """PER sample + update_priorities loop (C2 / C3 / per-rank shard of C5).
    sharded: None = follow the world size; False = every rank runs the un-sharded single-GPU step (the N=1
    equivalent measured in the same run, for C5's efficiency)."""
    torch, L, ffi = ctx.torch, ctx.L, ctx.ffi
    cap = 1_000_000
    R = sum(rows)
    h, cols, init_prio = make_prb(ctx, cap, rows, 7)
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
    sharded = ctx.world > 1 if sharded is None else sharded
    world_eff = ctx.world if sharded else 1
    stats = torch.empty(4, dtype=torch.float32, device="cuda")
    gstats = torch.empty((ctx.world, 4), dtype=torch.float32, device="cuda")
    exchange = os.environ.get("MRL_SHARD_EXCHANGE", "p2p")  # p2p | p2p_sample | nccl
    p2p = sharded and exchange in ("p2p", "p2p_sample") and ctx.world <= 32
    if p2p:  # peer mailboxes: the buffer's own kernels exchange the root statistics
        buf = C.create_string_buffer(64)
        ffi.check(L.mrl_prb_shard_init(h, ctx.rank, ctx.world, buf))
        ffi.check(L.mrl_prb_shard_set_mode(h, 0 if exchange == "p2p_sample" else 1))
        parts = [None] * ctx.world
        ctx.dist.all_gather_object(parts, buf.raw)
        allh = b"".join(parts)
        ffi.check(L.mrl_prb_shard_connect(h, C.create_string_buffer(allh, len(allh)), st))
        torch.cuda.synchronize()
        ctx.dist.barrier()

    def sample_into(u_ptr, idx_, w_, tab_):
        if p2p:
            ffi.check(L.mrl_prb_sample_sharded_p2p(h, batch, BETA, u_ptr, idx_.data_ptr(), w_.data_ptr(), tab_, st))
        elif sharded:  # root stats all-gather (16 B/rank over NVLink), then local strata
            ffi.check(L.mrl_prb_root_stats(h, stats.data_ptr(), st))
            ctx.dist.all_gather_into_tensor(gstats, stats)
            ffi.check(L.mrl_prb_sample_sharded(h, batch, BETA, u_ptr, gstats.data_ptr(), ctx.world,
                                               idx_.data_ptr(), w_.data_ptr(), tab_, st))
        else:
            ffi.check(L.mrl_prb_sample(h, batch, BETA, u_ptr, idx_.data_ptr(), w_.data_ptr(), tab_, st))

    check = None
    if parity and (sharded or ctx.world == 1):
        o, narrow = build_parity_oracle(ctx, cap, rows, init_prio, batch)
        check = parity_check(ctx, h, o, narrow, rows, batch, sample_into)
        del o
        if not check["ok"]:
            raise SystemExit(f"parity_check failed on rank {ctx.rank}: {check}")
    del init_prio

    def step(i):
        sample_into(None, idx, w, out_tab)
        ffi.check(L.mrl_prb_update(h, idx.data_ptr(), prio_ptr[i % pool], batch, st))

    if p2p:
        ffi.check(L.mrl_prb_shard_diag(h, None, None, None, None, 1, st))  # reset the exchange counters
    free_running = None
    if p2p and exchange == "p2p" and os.environ.get("MRL_BENCH_WAIT", "1") == "1":
        fr_ms, _ = ctx.timed(step, steps, warmup)
        v = [C.c_int64() for _ in range(4)]
        ffi.check(L.mrl_prb_shard_diag(h, *[C.byref(x) for x in v], 1, st))
        mhz = ctx.clocks.max_mhz or 1965.0
        free_running = {"value": world_eff * batch * steps / (fr_ms * 1e-3), "ms_per_step": fr_ms / steps,
                        "exchange_wait_us": ctx.max_over_ranks(v[0].value / max(v[1].value, 1) / mhz),
                        "calls_that_waited": v[2].value,
                        "note": "the same loop without the untimed mrl_prb_shard_wait before each timed step: flush-time "
                                "differences between the GPUs show up as waits in the timed sample of the faster ranks"}
        ctx.align = h
    total_ms, launches = ctx.timed(step, steps, warmup, sample_clocks=primary)
    ctx.align = None
    value = world_eff * batch * steps / (total_ms * 1e-3)
    exchange_info = None
    if p2p:
        v = [C.c_int64() for _ in range(4)]
        ffi.check(L.mrl_prb_shard_diag(h, *[C.byref(x) for x in v], 0, st))
        cyc, polled, waited, timeouts = (x.value for x in v)
        mhz = (ctx.clocks.summary().get("sm_mhz") or ctx.clocks.max_mhz or 1965.0)
        mine = cyc / max(polled, 1) / mhz  # us per sample call spent in the mailbox poll (warp 0 of CTA 0)
        worst = ctx.max_over_ranks(mine)
        exchange_info = {"protocol": "publish_on_mutation" if exchange == "p2p" else "publish_on_sample",
                         "exchange_wait_us": worst, "exchange_wait_us_rank0": mine, "polled_calls": polled,
                         "calls_that_waited": waited, "timeouts": timeouts,
                         "note": "SM cycles between entering and leaving the mailbox poll, warp 0 of CTA 0, per "
                                 "sample call (warm-up included), at the sampled SM clock; max over ranks"}
    if not details:
        ffi.check(L.mrl_prb_destroy(h))
        return {"workload": name, "metric": "per_sample_update_samples_per_s", "value": value, "unit": "samples/s",
                "ms_per_step": total_ms / steps, "steps": steps, "gpu_launches": launches,
                "parity_check": check, "exchange": exchange_info, "free_running": free_running}

    # ---- the same step captured once in a CUDA graph and replayed (what a launch-bound learner loop should do) ---
    # sample takes its stratum draws from a device array (the internal generator's counter advances on the host and
    # cannot be captured) and update its priorities from another: both are refreshed outside the timed region, like
    # every other input.  One graph launch instead of two kernel launches per step.
    graph_replay = None
    if not sharded and R <= 512:
        try:
            u_dev = torch.rand(batch, device="cuda", generator=g)
            p_dev = prio[0].clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                sst = side.cuda_stream
                for _ in range(2):  # eager first: one-off function attributes are set outside the capture
                    ffi.check(L.mrl_prb_sample(h, batch, BETA, u_dev.data_ptr(), idx.data_ptr(), w.data_ptr(), out_tab, sst))
                    ffi.check(L.mrl_prb_update(h, idx.data_ptr(), p_dev.data_ptr(), batch, sst))
                torch.cuda.synchronize()
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg, stream=side):
                    ffi.check(L.mrl_prb_sample(h, batch, BETA, u_dev.data_ptr(), idx.data_ptr(), w.data_ptr(), out_tab, sst))
                    ffi.check(L.mrl_prb_update(h, idx.data_ptr(), p_dev.data_ptr(), batch, sst))
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()

            def step_graph(i):
                cg.replay()

            def refresh(i):
                u_dev.uniform_(generator=g)
                p_dev.copy_(prio[i % pool])

            # (ctx.timed flushes, then records the events around step(): the refresh rides in front of the flush)
            def timed_graph():
                evs = []
                for i in range(10 + steps):
                    refresh(i)
                    ctx.flush()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    step_graph(i)
                    e1.record()
                    if i >= 10:
                        evs.append((e0, e1))
                torch.cuda.synchronize()
                return sum(a.elapsed_time(b) for a, b in evs)

            g_ms = timed_graph()
            graph_replay = {"value": batch * steps / (g_ms * 1e-3), "ms_per_step": g_ms / steps,
                            "note": "sample(uniforms from a device array) + update_priorities captured once in a CUDA "
                                    "graph, one replay per step, same flush discipline; two kernels, one launch"}
        except Exception as e:  # the stream-launch figures above stand on their own
            graph_replay = {"error": str(e)[:200]}

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
        "prb_update_sub_kernel": {"ms": t_update / ksteps, "bytes": (24 * L_1M + 20) * batch},
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

        # what the step actually launches for these rows: descent + DMA ring in ONE kernel
        def only_sample_rows(i):
            ffi.check(L.mrl_prb_sample(h, batch, BETA, None, idx.data_ptr(), w.data_ptr(), out_tab, st))

        l0 = ffi.launch_count()
        t_fused, _ = ctx.timed(only_sample_rows, ksteps, 3)
        per_call = (ffi.launch_count() - l0) / (ksteps + 3)
        kern["prb_sample_gather_kernel(descent+weights+row gather)"] = {
            "ms": t_fused / ksteps, "bytes": (4 * L_1M + 16 + 2 * R) * batch, "launches_per_call": per_call}
    else:
        kern["prb_sample_kernel(descent+weights)"]["bytes"] += 0
    # the roofline is quoted for the bandwidth-bound kernel of the step when there is one (the row gather
    # of the Atari-shaped buffer), else for the slowest (latency-bound) tree kernel
    fused_key = "prb_sample_gather_kernel(descent+weights+row gather)"
    dom = fused_key if fused_key in kern and kern[fused_key].get("launches_per_call") == 1 else (
        "gather_bulk_kernel" if "gather_bulk_kernel" in kern else max(kern, key=lambda k: kern[k]["ms"]))
    ach = kern[dom]["bytes"] / (kern[dom]["ms"] * 1e-3) / 1e9
    tkey = ("c3:" if R > 512 else "c2:") + {"prb_sample_kernel(descent+weights)": "prb_sample_kernel",
                                             "gather_bulk_kernel": "gather_bulk_kernel<4>",
                                             fused_key: "prb_sample_gather_kernel<4>"}.get(dom, dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s",
                "frac": ach / hbm, "traffic": ncu_traffic(tkey), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kern[dom]["bytes"],
                "kernel_ms": {k: v["ms"] for k, v in kern.items()},
                "ncu": {k: ncu_latency_evidence(("c3:" if R > 512 else "c2:") + k)
                        for k in (("prb_sample_gather_kernel<4>" if R > 512 else "prb_sample_kernel"), "prb_update_sub_kernel")},
                "note": ("tree kernels are dependent-load latency bound (L2/HBM round trips), "
                         "not bandwidth bound" if R <= 512 else
                         "row gather is HBM-bandwidth bound; since round 2 the descent and the DMA ring are one kernel "
                         "(kernel_ms also lists the former sample and gather kernels timed alone)")}

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

    # (the end-to-end leg always takes enough steps to be past its first-call effects -- staging buffers, mapped
    # memory, the event pool: 20 cold steps read 15 % low -- and says how many in e2e.steps)
    esteps = max(50, min(steps, 200)) if R > 512 else max(200, min(steps, 1000))
    e_ms = ctx.timed_host(step_host, esteps, 20)
    e2e = {"value": world_eff * batch * esteps / (e_ms * 1e-3), "unit": "samples/s",
           "h2d_bytes_per_step": batch * 12, "d2h_bytes_per_step": batch * (12 + R),
           "steps": esteps, "ms_per_step": e_ms / esteps}
    ffi.check(L.mrl_prb_destroy(h))
    return {"workload": name, "metric": "per_sample_update_samples_per_s", "value": value,
            "unit": "samples/s", "ms_per_step": total_ms / steps, "steps": steps,
            "gpu_launches": launches, "roofline": roofline, "e2e": e2e,
            "algorithmic_bytes_per_sample": 2 * R + 28 * L_1M + 36, "parity_check": check,
            "exchange": exchange_info, "free_running": free_running, "graph_replay": graph_replay}


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


def cpu_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_urb(seconds=2.0):
    """C1's CPU leg: the oracle's append 1 + sample 64 loop timed INSIDE the C library (no Python wrapper in the
    timed region), on one core and with the gather half split over all cores (as upstream's CPU kernel does)"""
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    cap, batch = 100_000, 64
    shapes, types = [(r,) for r in C2_ROWS], [np.uint8] * len(C2_ROWS)
    b = O.UniformBufferOracle(batch, cap, shapes, types, seed=5)
    b.insert([rng.integers(0, 256, size=(cap, r), dtype=np.uint8) for r in C2_ROWS])
    pool = [rng.integers(0, 256, size=(64, r), dtype=np.uint8) for r in C2_ROWS]
    b.bench_loop(2000, pool)
    per = b.bench_loop(20000, pool) / 20000
    steps = max(1000, int(seconds / max(per, 1e-9)))
    dt = b.bench_loop(steps, pool)
    nt = cpu_threads()
    mt_steps = max(200, min(steps, 20000))
    b.bench_loop(200, pool, threads=nt)
    dt_mt = b.bench_loop(mt_steps, pool, threads=nt)
    return {"value": batch * steps / dt, "unit": "samples/s", "cores": 1, "kind": "port",
            "sample": f"{steps} steps of insert 1 + sample 64 in {dt:.2f}s, timed inside the C oracle (orc_urb_bench_loop)",
            "all_cores": {"value": batch * mt_steps / dt_mt, "cores": nt,
                          "sample": f"{mt_steps} steps, the gather of every sample split over {nt} threads "
                                    "(orc_urb_gather_mt; a thread pool start per 5.6 KB batch costs more than it saves)"},
            "host": host_info()}


def cpu_wide_gather(seconds=3.0):
    """C3's row gather on the host: 512 Atari-shaped rows (2 x 28 224 B) per batch out of a 16 384-row buffer, one core
    and all cores (upstream splits this loop over its thread pool)"""
    from oracle import oracle as O
    rng = np.random.default_rng(2)
    cap, batch = 16384, 512
    rows = [28224, 28224]
    b = O.UniformBufferOracle(batch, cap, [(r,) for r in rows], [np.uint8] * 2, seed=5)
    block = [rng.integers(0, 256, size=(2048, r), dtype=np.uint8) for r in rows]
    for _ in range(cap // 2048):
        b.insert(block)
    slots = rng.integers(0, cap, size=batch)
    out = {}
    for name, nt in (("one_core", 1), ("all_cores", cpu_threads())):
        b.gather_bench(slots, nt, 3)
        per = b.gather_bench(slots, nt, 10) / 10
        reps = max(10, int(seconds / 2 / max(per, 1e-9)))
        dt = b.gather_bench(slots, nt, reps)
        out[name] = {"cores": nt, "GBps": 2 * sum(rows) * batch * reps / dt / 1e9, "ms_per_batch": dt / reps * 1e3,
                     "reps": reps}
    return out


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

    # GAE + PPO's advantage normalisation as one operation (moments fused into the scan's store phase) against
    # the two-step form (scan, then mrl_normalize = moments pass + apply pass)
    def gae_norm_fused(i):
        ffi.check(L.mrl_gae_normalized(r.data_ptr(), v.data_ptr(), None, v[T].data_ptr(), d.data_ptr(), adv.data_ptr(),
                                       ret.data_ptr(), T, B, 0.99, 0.95, 1e-8, 0, 0, st))

    def gae_then_norm(i):
        gae_tm(i)
        ffi.check(L.mrl_normalize(adv.data_ptr(), adv.data_ptr(), n, 1e-8, st))

    run("gae+normalize_fused_moments[T,B]", gae_norm_fused, 17 + 8, 4 * B)
    run("gae_then_normalize_separate[T,B]", gae_then_norm, 17 + 12, 4 * B)

    # element-wise error of the PARALLEL kernels against the sequential kernel (which is bit-identical to the CPU
    # oracle: tests/test_gpu_scans.py).  north_star words the tolerance as "1e-5 relative"; sums that cancel to ~0
    # have no meaningful element-wise relative error (the fp32 oracle itself misses 1e-5 against float64 there),
    # so the bound held by the tests is |gpu - oracle| <= 1e-5 * max(||oracle||_inf, 1) and the distribution of the
    # element-wise relative error is reported here.
    def err_stats(par, seq):
        diff = (par.double() - seq.double()).abs()
        rel = diff / seq.double().abs().clamp_min(1e-30)
        srt = rel.flatten().sort().values
        q = lambda f: float(srt[min(srt.numel() - 1, int(f * srt.numel()))])  # noqa: E731
        scale = max(1.0, float(seq.abs().max()))
        return {"p50": q(0.5), "p99": q(0.99), "p999": q(0.999), "max_rel": float(srt[-1]), "max_abs": float(diff.max()),
                "max_abs_over_norm": float(diff.max()) / scale, "frac_elems_rel_gt_1e-5": float((rel > 1e-5).double().mean()),
                "frac_elems_rel_gt_1e-5_and_abs_gt_1e-5": float(((rel > 1e-5) & (diff > 1e-5)).double().mean())}

    seq_adv, seq_ret = torch.empty_like(adv), torch.empty_like(ret)
    ffi.check(L.mrl_gae(r.data_ptr(), v.data_ptr(), None, v[T].data_ptr(), d.data_ptr(), seq_adv.data_ptr(),
                        seq_ret.data_ptr(), T, B, 0.99, 0.95, 0, 1, st))
    gae_tm(0)
    scan_error = {"gae_time_major_adv": err_stats(adv, seq_adv), "gae_time_major_ret": err_stats(ret, seq_ret)}
    ffi.check(L.mrl_gae(r.data_ptr(), v.data_ptr(), None, v[T].data_ptr(), d.data_ptr(), seq_adv.data_ptr(),
                        seq_ret.data_ptr(), T, B, 0.99, 0.95, 1, 1, st))
    gae_bm(0)
    scan_error["gae_batch_major_adv"] = err_stats(adv, seq_adv)
    ffi.check(L.mrl_discounted_return(r.data_ptr(), d.data_ptr(), v[T].data_ptr(), seq_adv.data_ptr(), T, B, 1, 0.99, 0, 1, st))
    dr_tm(0)
    scan_error["discounted_return_time_major"] = err_stats(adv, seq_adv)
    scan_error["tolerance_held_by_the_tests"] = "|gpu - oracle| <= 1e-5 * max(||oracle||_inf, 1) (norm-wise); north_star: 1e-5 relative"
    scan_error["reference"] = "sequential GPU kernel = CPU oracle bit for bit (tests/test_gpu_scans.py)"
    del seq_adv, seq_ret
    main = out["gae_time_major[T,B]"]
    # quoted fraction = launches back to back over five rotated buffer sets (every launch also carries its
    # predecessor's write-back): the conservative one.  After a flush the kernel's own write-back partly outlives it
    # (ncu counts 94 of 143 MB of DRAM traffic inside the kernel), which flatters the per-launch figure: after_flush.
    roofline = {"bound": "hbm", "kernel": "scan_tm_strip_kernel<GAE, done, 32 cols, 16 steps/lane>",
                "achieved": main["back_to_back_GBps"], "peak": hbm, "unit": "GB/s", "frac": main["back_to_back_frac"],
                "after_flush": {"achieved": main["GBps"], "frac": main["frac"],
                                "note": "CUDA events around one launch after an L2 flush (the discipline of `value`)"},
                "traffic": ncu_traffic("c4:scan_tm_strip_kernel<1, 1, 32, 16, 1, 0>"),
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
    roofline["batch_major_kernel"] = {"kernel": "scan_bm_row_kernel<GAE, done>", "achieved": bm["back_to_back_GBps"],
                                      "frac": bm["back_to_back_frac"], "after_flush_frac": bm["frac"],
                                      "traffic": ncu_traffic("c4:scan_bm_row_kernel<1, 1, 0, 0, 8>")}
    roofline["discounted_return"] = {
        "time_major": {"kernel": "scan_tm_strip_kernel<DR, done, 32 cols, 32 steps/lane>",
                       "frac": out["discounted_return_time_major"]["back_to_back_frac"],
                       "after_flush_frac": out["discounted_return_time_major"]["frac"]},
        "batch_major": {"kernel": "scan_bm_row_kernel<DR, done>",
                        "frac": out["discounted_return_batch_major"]["back_to_back_frac"],
                        "after_flush_frac": out["discounted_return_batch_major"]["frac"]},
        "note": "9 B per element; north_star's 60 % of the nominal 8 TB/s is 0.73 of the measured peak: not reached"}
    roofline["note"] = "frac = 5 rotated buffer sets, launches back to back in one event pair; after_flush = per launch after an L2 flush"
    w = ctx.world
    e2e["value"] *= w
    name = "C4 GAE 4096 envs x 2048 steps fp32 gamma=0.99 lambda=0.95 done~Bernoulli(1/200)"
    if w > 1:
        name += f", env axis sharded over {w} ranks (4096 envs per rank, no communication; whole-job elements/s)"
    return {"workload": name,
            "metric": "gae_elements_per_s", "value": main["elems_per_s"] * w, "unit": "elements/s",
            "ms_per_step": main["ms"], "steps": steps, "roofline": roofline, "e2e": e2e,
            "scan_error": scan_error, "variants": out}


# ---- CPU baselines (oracle port; the only place bench.py executes oracle/) ---------------
def cpu_per(rows, batch, capacity, seconds=8.0, max_steps=200000, warm_seconds=0.5):
    """PER sample + update_priorities on the CPU oracle, the loop timed inside the C library"""
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    shapes = [(r,) for r in rows]
    types = [np.uint8] * len(rows)
    b = O.PriorityBufferOracle(ALPHA, capacity, batch, shapes, types, 7, 0)
    chunk = max(1, min(capacity, (64 << 20) // max(sum(rows), 1)))
    block = [rng.integers(0, 256, size=(chunk, r), dtype=np.uint8) for r in rows]
    for s_ in range(0, capacity, chunk):
        n = min(chunk, capacity - s_)
        b.insert_batch([blk[:n] for blk in block])
    pr = (np.abs(rng.normal(size=capacity)) + 1e-3).astype(np.float32)
    b.update_priorities(np.arange(capacity), pr)
    prio = (np.abs(rng.normal(size=(64, batch))) + 1e-3).astype(np.float32)
    per = b.bench_loop(20, BETA, prio) / 20
    b.bench_loop(max(3, int(warm_seconds / max(per, 1e-9))), BETA, prio)  # time-based warm-up (caches, page faults)
    steps = max(20, min(max_steps, int(seconds / max(per, 1e-9))))
    dt = b.bench_loop(steps, BETA, prio)
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


C2_NAME = ("C2 PER cap 1M (2^20 leaves) 40-byte rows batch 256 alpha 0.6 beta 0.4, "
           "sample+update_priorities per step")


def headline_config(world):
    """the `config` object of the JSON line: the same dict in both arms (the driver compares them key by key)"""
    return {"workload": C2_NAME, "capacity": 1_000_000, "batch": 256, "row_bytes": sum(C2_ROWS), "alpha": ALPHA,
            "beta": BETA, "per_rank": True,
            "l2": "flushed before every timed iteration (256 MiB written, then 256 MiB read; untimed); step time = "
                  "sum of per-iteration CUDA-event durations on the launching stream (reference arm: host clock "
                  "around the C loop)",
            "sharding": ("per-rank sub-buffers (1M rows each); root statistics published to peer mailboxes over NVLink "
                         "by the mutating kernels (MRL_SHARD_EXCHANGE=p2p_sample: by the sample kernel; nccl: NCCL "
                         "all-gather)" if world > 1 else "single GPU")}


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
    # warm-up: the W steps asked for AND at least half a second, so that a 20-step run is not timed cold
    per = b.bench_loop(max(args.warmup, 3), BETA, prio) / max(args.warmup, 3)
    warm = max(args.warmup, 3) + max(0, int(0.5 / max(per, 1e-9)))
    b.bench_loop(warm - max(args.warmup, 3), BETA, prio) if warm > max(args.warmup, 3) else None
    # each "step" of this arm is a bounded sample of the workload: K_inner back-to-back sample+update steps
    inner = max(1, int(1.0 / max(per, 1e-9) / max(args.steps, 1)))  # ~1 s of CPU work over the whole run
    dt = b.bench_loop(args.steps * inner, BETA, prio)
    value = batch * args.steps * inner / dt
    line = {
        "impl": "reference", "metric": "per_sample_update_samples_per_s", "value": value,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dt * 1e3 / (args.steps * inner), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": 1, "kind": "port",
                         "sample": f"{args.steps} x {inner} steps of sample(256)+update(256) on the full 1M-leaf tree after "
                                   f"{warm} warm-up steps, timed inside the C oracle (orc_prb_bench_loop); MindSpore (the "
                                   "reference's kernels) is not installable here, so this is the oracle port of its "
                                   "sequential CPU algorithm (one core: upstream's PER kernels are sequential loops)",
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
    ap.add_argument("--only", default="all", choices=["all", "c1", "c2", "c3", "c4", "c5", "insert"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baselines")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args, emit)
        return

    ctx = Ctx(args)
    primary, also, c5 = None, [], None
    C5_NAME = "C5 sharded PER 1M rows/rank, 40-byte rows, batch 4096/rank, sample+update_priorities per step"
    if args.only in ("all", "c2"):
        primary = bench_per(ctx, C2_NAME, C2_ROWS, 256, args.steps, args.warmup, True)
    if args.only in ("all", "c5"):
        # C5 at EVERY N (N = 1: one shard, batch 4096), so that the records hold a 1 -> 8 curve for it; at N > 1
        # every rank also runs the un-sharded step on its own GPU in the same run (the N = 1 equivalent)
        c5_steps = max(20, min(args.steps, 300))
        r = bench_per(ctx, C5_NAME, C2_ROWS, 4096, c5_steps, args.warmup, primary is None)
        c5 = {"workload": C5_NAME, "metric": r["metric"], "unit": r["unit"], "value": r["value"], "n_gpus": ctx.world,
              "ms_per_step": r["ms_per_step"], "steps": r["steps"], "e2e": r["e2e"]["value"],
              "parity_check": r["parity_check"], "exchange": r["exchange"], "free_running": r["free_running"]}
        if ctx.world > 1:
            solo = bench_per(ctx, C5_NAME + " (un-sharded, every rank alone)", C2_ROWS, 4096, c5_steps, args.warmup,
                             False, sharded=False, parity=False, details=False)
            c5["single_gpu_same_run"] = {"value": solo["value"], "ms_per_step": solo["ms_per_step"],
                                         "note": "the same step without the exchange, all ranks at once, max over ranks"}
            c5["weak_scaling_efficiency_vs_same_run_single_gpu"] = r["value"] / (ctx.world * solo["value"])
        (also.append(r) if primary else None)
        primary = primary or r
    if ctx.world == 1 and args.only in ("all", "c3"):
        r = bench_per(ctx, "C3 Atari-shaped PER cap 1M obs 84x84x4 u8 (x2) batch 512", C3_ROWS, 512,
                      max(20, min(args.steps, 200)), args.warmup, primary is None)
        (also.append(r) if primary else None)
        primary = primary or r
    if args.only in ("all", "c4") and (ctx.world == 1 or args.only == "all"):
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

    if ctx.rank == 0:
        if not args.no_cpu:
            for r in [primary] + also:
                wl = r["workload"]
                if wl.startswith("C2") or wl.startswith("C5"):
                    b = 256 if wl.startswith("C2") else 4096
                    v, n, dt = cpu_per(C2_ROWS, b, 1_000_000, seconds=8.0 if r is primary else 4.0)
                    r["cpu_baseline"] = {
                        "value": v, "unit": "samples/s", "cores": 1, "kind": "port",
                        "sample": f"{n} steps of sample({b})+update({b}) in {dt:.1f}s on the full 1M-leaf oracle tree after a "
                                  "0.5 s warm-up, timed inside the C oracle (one core: upstream's PER kernels are "
                                  "sequential loops)", "host": host_info()}
                elif wl.startswith("C3"):
                    v, n, dt = cpu_per(C3_ROWS, 512, 32768, seconds=6.0)
                    r["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": 1, "kind": "port",
                                         "sample": f"{n} steps in {dt:.1f}s on a 32768-row oracle buffer (1.8 GB of rows), "
                                                   "timed inside the C oracle",
                                         "row_gather_on_host": cpu_wide_gather(), "host": host_info()}
                elif wl.startswith("C1"):
                    r["cpu_baseline"] = cpu_urb()
                elif wl.startswith("C4"):
                    g, npl = cpu_gae(2048, 4096)
                    r["cpu_baseline"] = {"value": g, "unit": "elements/s", "cores": 1, "kind": "port",
                                         "sample": "full 2048x4096 GAE on the C oracle, best of 3",
                                         "numpy_loop_discounted_return_elements_per_s": npl, "host": host_info()}
        cfg = headline_config(ctx.world)
        cfg["workload"] = primary["workload"]
        if ctx.world > 1 and os.environ.get("MRL_BENCH_WAIT", "1") == "1":
            cfg["l2"] += ("; at N > 1 every rank also waits, untimed, until its peers' previous publications have arrived "
                          "(mrl_prb_shard_wait) before a timed step starts; the loop without that wait is reported as "
                          "free_running")
        line = {
            "metric": primary["metric"], "value": primary["value"], "unit": primary["unit"],
            "n_gpus": ctx.world, "steps": primary["steps"], "warmup": args.warmup,
            "ms_per_step": primary["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "clocks": ctx.clocks.summary(),
            "e2e": primary["e2e"], "gpu_launches": primary.get("gpu_launches"),
            "roofline": primary.get("roofline"), "cpu_baseline": primary.get("cpu_baseline"),
            "parity_check": primary.get("parity_check"), "exchange": primary.get("exchange"),
            "free_running": primary.get("free_running"), "graph_replay": primary.get("graph_replay"),
            "algorithmic_bytes_per_sample": primary.get("algorithmic_bytes_per_sample"),
            "variants": primary.get("variants"), "scan_error": primary.get("scan_error"),
            "c5": c5,
            "also": also,
        }
        emit(line)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
