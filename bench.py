#!/usr/bin/env python
"""bench.py -- the measurement contract for the B200 KV-cache hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload all|c3|c4|c5|paged|c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[2], the configuration the metric "KV quant/dequant HBM GB/s at
1/2/4/8 B200" is quoted on; it fits one GPU):
    Llama-3-8B-shape synthetic KV cache -- 32 layers x {K, V}, each [B=16, H=8, T=32768, D=128] bf16
    (1 GiB) -- INT4 pack / unpack with a per-tensor scale, head-sharded over the N ranks.
One step = one pass over the whole batch: for every layer tensor  absmax (phase 1)  ->  [N > 1: ONE
allreduce(MAX) of the 64 absmax words over NCCL]  ->  quantize + nibble-pack (phase 2)  ->  unpack +
dequantize to bf16.  value = algorithmic bytes of the whole job / step time (max over ranks), inputs
resident in HBM (64 GiB >> 126 MB L2, so no flush is needed).  e2e = the same metric through the public
Python API (kvq_b200.quantize_int4_per_tensor_packed / dequantize_int4_per_tensor_packed) with HOST
buffers: pinned host -> device copy of every input and device -> host copy of the packed codes, the scale
and the dequantized tensor inside the timed region.

The same JSON line carries one object per other BASELINE config / kernel family, each with its own
`roofline` object (algorithmic bytes from BASELINE.md's table, CUDA-event time, fraction of the measured
copy peak) and clock sample:
    per_token : per-token (cache granularity) fused quantize INT4 / INT8 and per-token dequantize, C3 shape
    c4        : BASELINE configs[3] -- strided / block / budget / prefix+window compaction and chunk-summary
                pooling at T = 131072, L = 32 x {K, V} in ONE launch each, head-sharded over the N ranks
    decode    : BASELINE configs[4] -- fused dequant decode attention, 80 layers, B8 x 64k, INT8-K/INT4-V,
                MHA shape and GQA-8, contiguous and paged stores; at N > 1 also the head-sharded decode append
                (decode.append_sharded: exchange inside the quantize kernel, eager and in a CUDA graph, vs the NCCL form)
    paged     : PagedKVCache append / stitch kernels
    c2        : BASELINE configs[1] (and [0] on the GPU) -- GPT-2 small tok/s through KVCacheBenchmarker.benchmark_method
--workload picks a subset of the legs (profiling runs); the default runs them all.

--impl reference times the reference's CPU path (the oracle port of its torch implementation, all host
cores) on a bounded sample of the same workload; rank 0 only.
"""

from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import statistics
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "efficient-llm-inference_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

_JSON_FD = 1


def emit(line: dict) -> None:
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


METRIC = "kv_quant_dequant_hbm_gbs"
UNIT = "GB/s"
C3 = dict(L=32, B=16, H=8, T=32768, D=128)
C4 = dict(L=32, B=1, H=8, T=131072, D=128)
C5 = dict(L=80, B=8, H=8, T=65536, D=128)  # Llama-3-70B-shape KV (BASELINE.json configs[4]), INT8-K / INT4-V
LEGS = ("c3", "per_token", "c4", "c5", "paged", "e2e", "c2", "cpu")


# ----------------------------------------------------------------------------- helpers


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy_)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(kernel: str):
    """dram bytes (read+write) per launch of `kernel` from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def roofline(kernel: str, alg_bytes: float, ms: float, launches_timed: int, traffic=None, note=None) -> dict:
    """HBM roofline object of one kernel: algorithmic bytes per launch / CUDA-event duration vs the measured copy peak."""
    peak, src = measured_peak()
    achieved = alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    r = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "frac_of_8TBs": achieved / 8000.0, "peak_source": src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": ms,
         "launches_timed": launches_timed, "traffic": traffic}
    if note:
        r["note"] = note
    return r


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            pass
    return local_rank


class Ctx:
    """Per-process state shared by the legs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from kvq_b200 import _lib as L

        self.args = args
        self.rank, self.local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world != args.gpus:
            if self.world == 1 and args.gpus > 1:
                raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
            args.gpus = self.world
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.L, self.lib, self.torch, self.dist = L, L.lib(), torch, dist
        sm = ctypes.c_int()
        L.check(self.lib.kvq_device_info(ctypes.byref(sm), None, None), "kvq_device_info")
        self.sm_count = sm.value
        self.launches = 0
        self._flush = None

    def stream(self):
        return self.L.stream_ptr(self.dev)

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        if self.world > 1:
            # the hosts leave the barrier up to a millisecond apart; a stream-ordered (not host-synchronised) tiny allreduce makes the
            # DEVICE timelines start together: every rank's next event is recorded when this kernel ends, which it does on all
            # ranks at once.  Without it the early ranks' timed regions include the wait for the late ones' first step.
            if not hasattr(self, "_align"):
                self._align = self.torch.zeros(1, device=self.dev)
            self.dist.all_reduce(self._align)

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def sampler(self):
        s = ClockSampler(physical_gpu_index(self.local_rank))
        s.start()
        return s

    def flush_l2(self):
        """Write a buffer twice the size of the 126 MB L2 (between timed launches whose working set could stay resident)."""
        if self._flush is None:
            self._flush = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)
        self._flush.zero_()

    def time_launches(self, fn, reps: int, flush: bool = True, warm: int = 3):
        """Mean / min CUDA-event duration (ms) of fn() -- one launch (or one fixed group of launches) -- over `reps` runs on
        torch's current stream, L2 flushed before each run when `flush`; max over ranks of the mean."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.sync_all()
        evs = []
        for _ in range(reps):
            if flush:
                self.flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        self.sync_all()
        ts = [a.elapsed_time(b) for a, b in evs]
        return self.max_over_ranks(statistics.mean(ts)), self.max_over_ranks(min(ts))


# ----------------------------------------------------------------------------- the reference's own torch code (baseline/_ref)


def load_reference_package():
    """The UNMODIFIED reference package (git-ignored copy under baseline/_ref, placed by tools/install_reference.sh;
    /root/reference in the build container) under the alias `kvq_refsrc`, without running its heavy top-level __init__
    and without shadowing this repo's own `src` shim.  None when absent."""
    if "kvq_refsrc" in sys.modules:
        return sys.modules["kvq_refsrc"]
    for base in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        path = os.path.join(base, "src")
        if os.path.exists(os.path.join(path, "quantization", "ops.py")):
            sys.dont_write_bytecode = True
            pkg = types.ModuleType("kvq_refsrc")
            pkg.__path__ = [path]
            sys.modules["kvq_refsrc"] = pkg
            return pkg
    return None


def reference_torch_sample(cfg, device: str, seconds: float, T_slice: int):
    """The reference's own functions (src/quantization/ops.py:33-65, :93-133, unmodified) on a T-slice of the C3 tensors:
    quantize_int4_per_tensor_packed + dequantize_int4_per_tensor_packed, torch ops on `device`.  GB/s on the same
    7 B/element algorithmic bytes as the headline.  None when the reference package is not importable."""
    import torch

    if load_reference_package() is None:
        return None
    try:
        ops = importlib.import_module("kvq_refsrc.quantization.ops")
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"{type(e).__name__}: {e}"}
    B, H, D = cfg["B"], cfg["H"], cfg["D"]
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, H, T_slice, D, generator=g, dtype=torch.float32).to(torch.bfloat16).to(device)
    cores = os.cpu_count() or 1
    if device == "cpu":
        torch.set_num_threads(cores)

    def one():
        p, s, last = ops.quantize_int4_per_tensor_packed(x)
        return ops.dequantize_int4_per_tensor_packed(p, s, last, torch.bfloat16)

    one()
    if device != "cpu":
        torch.cuda.synchronize()
    passes, t0 = 0, time.perf_counter()
    while True:
        one()
        if device != "cpu":
            torch.cuda.synchronize()
        passes += 1
        el = time.perf_counter() - t0
        if el >= seconds or passes >= 4096:
            break
    gbs = 7.0 * x.numel() * passes / el / 1e9
    return {"value": gbs, "unit": UNIT, "cores": cores if device == "cpu" else None, "kind": "reference", "device": device,
            "sample": (f"{passes} passes of the reference's quantize_int4_per_tensor_packed + dequantize_int4_per_tensor_packed "
                       f"(torch ops, unmodified src/quantization/ops.py) over one [B={B},H={H},T={T_slice},D={D}] bf16 tensor on "
                       f"{device}, {el:.1f} s")}


# ----------------------------------------------------------------------------- CPU reference arm


def cpu_reference_sample(cfg, seconds=12.0, T_slice=4096):
    """The reference's CPU path (oracle port, OpenMP over all host cores) on a T-slice of the same tensors:
    absmax -> INT4 quantize+pack -> unpack+dequantize to bf16, per tensor.  Returns (GB/s, cores, sample text, passes)."""
    import torch
    from oracle import kvq_oracle as o

    o.build()
    cores = os.cpu_count() or 1
    o.set_num_threads(cores)
    B, H, D = cfg["B"], cfg["H"], cfg["D"]
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, H, T_slice, D, generator=g, dtype=torch.float32).to(torch.bfloat16)
    n = x.numel()
    nbytes = 7 * n  # 2N absmax + 2.5N quant + 2.5N unpack
    o.quantize_int4_per_tensor_packed(x)  # warm (page faults, thread pool)
    passes, t0 = 0, time.perf_counter()
    while True:
        p, s, last = o.quantize_int4_per_tensor_packed(x)
        o.dequantize_int4_per_tensor_packed(p, s, last, torch.bfloat16)
        passes += 1
        el = time.perf_counter() - t0
        if el >= seconds or passes >= 4096:
            break
    gbs = nbytes * passes / el / 1e9
    sample = (f"{passes} passes over one [B={B},H={H},T={T_slice},D={D}] bf16 tensor (T sliced from {cfg['T']}); absmax + INT4 "
              f"quantize/pack + unpack/dequantize by the C oracle (OpenMP, {cores} threads), {el:.1f} s")
    return gbs, cores, sample, passes


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 3.0  # seconds of CPU work per "step" (a bounded sample of the workload)
    for _ in range(max(args.warmup, 0)):
        cpu_reference_sample(cfg, seconds=0.5)
    vals, t_all = [], time.perf_counter()
    sample = ""
    for _ in range(max(args.steps, 1)):
        gbs, cores, sample, _ = cpu_reference_sample(cfg, seconds=per_step)
        vals.append(gbs)
    value = statistics.mean(vals)
    ms = (time.perf_counter() - t_all) * 1e3 / max(args.steps, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(cfg, args.gpus, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": "per step: " + sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(cfg, n_gpus, where):
    return {
        "workload": (f"Llama-3-8B-shape synthetic KV (L={cfg['L']} x {{K,V}}, B={cfg['B']}, H={cfg['H']} KV heads, T={cfg['T']}, "
                     f"D={cfg['D']}, bf16): INT4 pack/unpack, per-tensor scale, head-sharded over {n_gpus} GPU(s)"),
        "baseline_config": "BASELINE.json configs[2]",
        "step": ("absmax of all 64 layer tensors (1 launch) -> quantize+pack (1 launch) -> unpack+dequantize (1 launch)" if n_gpus == 1 or where != "gpu" else
                 "absmax (1 launch) -> max over ranks of the 64 absmax words (see `collective`) -> quantize+pack (1 launch) -> unpack+dequantize (1 launch)"),
        "algorithmic_bytes_per_element": 7.0,
        "l2_policy": "inputs (64 GiB whole job) are far larger than the 126 MB L2; no flush needed (the small-output legs flush L2 before every timed launch)",
        "parallelism": f"kv-head shard x{n_gpus}" if where == "gpu" else "cpu",
    }


# ----------------------------------------------------------------------------- GPU arm: C3 (+ per-token leg on the same buffers)


def run_c3(cx: Ctx, cfg, legs):
    torch, dist, L, lib, dev, rank, world, args = cx.torch, cx.dist, cx.L, cx.lib, cx.dev, cx.rank, cx.world, cx.args
    from kvq_b200.distributed import head_shard_range
    from kvq_b200.quantization import TensorTable

    Lyr, B, H, T, D = cfg["L"], cfg["B"], cfg["H"], cfg["T"], cfg["D"]
    lo, hi = head_shard_range(H, rank, world)
    Hl = hi - lo
    n_tensors = 2 * Lyr
    n_local = B * Hl * T * D
    n_global = B * H * T * D

    # synthetic inputs: one generator per (layer, K|V, head) so that 1-GPU and N-GPU data agree
    xs = []
    for layer in range(Lyr):
        for kv in range(2):
            x = torch.empty(B, Hl, T, D, device=dev, dtype=torch.bfloat16)
            for h in range(lo, hi):
                g = torch.Generator(device=dev).manual_seed(1234 + 1000 * layer + kv + 100000 * h)
                x[:, h - lo] = torch.randn(B, T, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
            xs.append(x)
    packed = [torch.empty(B, Hl, T, D // 2, device=dev, dtype=torch.uint8) for _ in range(n_tensors)]
    outs = [torch.empty(B, Hl, T, D, device=dev, dtype=torch.bfloat16) for _ in range(n_tensors)]
    bits = torch.zeros(n_tensors, dtype=torch.int32, device=dev)
    scales = torch.empty(n_tensors, dtype=torch.bfloat16, device=dev)
    tx, tp, to = TensorTable(xs), TensorTable(packed), TensorTable(outs)
    stream = cx.stream()
    dom_events = []  # (start, end) around every launch of the dominant kernel inside the timed region
    front_events = []  # (start, end) around absmax -> exchange -> quantize of every timed step
    half = n_tensors // 2
    from kvq_b200.distributed import AbsmaxExchange
    xch = AbsmaxExchange(dev) if world > 1 and not args.nccl_absmax else None
    use_xch = xch is not None and xch.available
    use_fused = use_xch and not args.side_stream_exchange
    bits_local = torch.zeros(n_tensors, dtype=torch.int32, device=dev)
    if use_xch and not use_fused:
        side = torch.cuda.Stream(dev, priority=-1)
        ev_abs, ev_x = [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)]
        bits_halves = [bits[:half], bits[half:]]

    def step(timed: bool):
        # one launch per phase for the whole batch of 64 layer tensors (kvq_*_batched).  Head-sharded (N > 1) the max over ranks of
        # the 64 absmax words sits between absmax and quantize: by default INSIDE the two kernels (publish / gather over peer-mapped
        # memory); as A/B a stand-alone exchange kernel (--side-stream-exchange) or NCCL's allreduce (--nccl-absmax), both with the
        # batch in two halves so that the exchange of one half overlaps the other half's kernel
        n_launch = 3
        if timed:
            e_in = torch.cuda.Event(enable_timing=True)
            e_in.record()
        if use_fused:
            # the exchange lives INSIDE the two kernels: the absmax kernel's last CTA per tensor stores the word into every rank's
            # peer-mapped buffer, the quantize kernel's CTAs gather the world's words of their tensor in the prologue.  Three
            # launches per step like the 1-GPU step, one stream, no collective call.
            seq = xch.next_fused_seq()
            L.check(lib.kvq_absmax_batched_publish(tx.ptrs.data_ptr(), L.BF16, n_tensors, n_local, bits_local.data_ptr(), xch.peer_table_ptr,
                                                   xch.rank, xch.world, seq, stream), "kvq_absmax_batched_publish")
            L.check(lib.kvq_quant_batched_gather(tx.ptrs.data_ptr(), L.BF16, n_tensors, n_local, 4 | L.ABSMAX_COVERS_INPUT, bits.data_ptr(), 1e-8,
                                                 tp.ptrs.data_ptr(), scales.data_ptr(), None, xch.local_buffer_ptr, xch.world, seq, stream),
                    "kvq_quant_batched_gather")
        elif use_xch:
            # A/B (--side-stream-exchange): the exchange kernel (peer-mapped memory, no NCCL call on the step) runs on a high-priority side stream: the batch
            # goes through absmax / quantize in two halves so that the wait for the slowest rank's words of one half hides
            # behind the other half's kernel
            main = torch.cuda.current_stream(dev)
            for i, (lo_t, cnt) in enumerate(((0, half), (half, n_tensors - half))):
                L.check(lib.kvq_absmax_batched(tx.ptrs.data_ptr() + 8 * lo_t, L.BF16, cnt, n_local, bits.data_ptr() + 4 * lo_t, stream),
                        "kvq_absmax_batched")
                ev_abs[i].record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ev_abs[i])
                    xch.exchange_(bits_halves[i])
                    ev_x[i].record(side)
            for i, (lo_t, cnt) in enumerate(((0, half), (half, n_tensors - half))):
                main.wait_event(ev_x[i])
                L.check(lib.kvq_quant_batched(tx.ptrs.data_ptr() + 8 * lo_t, L.BF16, cnt, n_local, 4 | L.ABSMAX_COVERS_INPUT,
                                              bits.data_ptr() + 4 * lo_t, 1e-8, tp.ptrs.data_ptr() + 8 * lo_t,
                                              scales.data_ptr() + 2 * lo_t, None, stream), "kvq_quant_batched")
            n_launch = 7
        elif world > 1:
            works = []
            for lo_t, cnt in ((0, half), (half, n_tensors - half)):
                L.check(lib.kvq_absmax_batched(tx.ptrs.data_ptr() + 8 * lo_t, L.BF16, cnt, n_local, bits.data_ptr() + 4 * lo_t, stream),
                        "kvq_absmax_batched")
                works.append(dist.all_reduce(bits[lo_t: lo_t + cnt], op=dist.ReduceOp.MAX, async_op=True))  # the only collective on the path
            for (lo_t, cnt), wk in zip(((0, half), (half, n_tensors - half)), works):
                wk.wait()  # stream-level wait
                L.check(lib.kvq_quant_batched(tx.ptrs.data_ptr() + 8 * lo_t, L.BF16, cnt, n_local, 4 | L.ABSMAX_COVERS_INPUT,
                                              bits.data_ptr() + 4 * lo_t, 1e-8, tp.ptrs.data_ptr() + 8 * lo_t,
                                              scales.data_ptr() + 2 * lo_t, None, stream), "kvq_quant_batched")
            n_launch = 5
        else:
            L.check(lib.kvq_absmax_batched(tx.ptrs.data_ptr(), L.BF16, n_tensors, n_local, bits.data_ptr(), stream), "kvq_absmax_batched")
            L.check(lib.kvq_quant_batched(tx.ptrs.data_ptr(), L.BF16, n_tensors, n_local, 4 | L.ABSMAX_COVERS_INPUT, bits.data_ptr(), 1e-8, tp.ptrs.data_ptr(),
                                          scales.data_ptr(), None, stream), "kvq_quant_batched")
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            front_events.append((e_in, e0))
        L.check(lib.kvq_dequant_batched(tp.ptrs.data_ptr(), 4, n_tensors, n_local, scales.data_ptr(), L.BF16, to.ptrs.data_ptr(),
                                        L.BF16, stream), "kvq_dequant_batched")
        if timed:
            e1.record()
            dom_events.append((e0, e1))
        return n_launch

    for _ in range(max(args.warmup, 3)):
        step(False)
    cx.sync_all()
    sampler = cx.sampler()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ev0.record()
    for _ in range(args.steps):
        launches += step(True)
    ev1.record()
    cx.sync_all()
    clocks = sampler.stop()
    my_ms = ev0.elapsed_time(ev1) / args.steps
    ms_step = cx.max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    # per-rank view (explains the scaling figure): every rank's own ms per step, and its absmax -> exchange -> quantize / dequantize split
    mine = torch.tensor([my_ms, statistics.mean(a.elapsed_time(b) for a, b in front_events),
                         statistics.mean(a.elapsed_time(b) for a, b in dom_events)], device=dev)
    if world > 1:
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    per_rank = {"ms_per_step": [round(float(t[0]), 4) for t in allr], "absmax_exchange_quant_ms": [round(float(t[1]), 4) for t in allr],
                "dequant_ms": [round(float(t[2]), 4) for t in allr]}
    job_bytes = 7.0 * n_global * n_tensors
    value = job_bytes / (ms_step * 1e-3) / 1e9
    cx.launches += launches

    # roofline of the dominant kernel (dequant_flat_kernel<bf16, INT4>): algorithmic N*(0.5+2) bytes per launch
    dom_ms = statistics.mean(a.elapsed_time(b) for a, b in dom_events)
    dom_bytes = 2.5 * n_local * n_tensors  # one launch covers all layer tensors
    tr = ncu_traffic("dequant_flat_int4_bf16")
    roof = roofline("dequant_flat_kernel<bf16,int4> (unpack+dequantize)", dom_bytes, dom_ms, len(dom_events),
                    None if tr is None else tr * n_tensors * n_local / (16 * 8 * 32768 * 128))

    # ---- parity inside the run.  (1) the allreduced absmax words against an INDEPENDENT global absmax: torch's own
    # abs().max() of the local shard, max-combined over ranks through a different path (all_gather + host max)
    own = torch.stack([x.abs().max().float() for x in xs[:4]])
    if world > 1:
        gathered = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(gathered, own)
        own = torch.stack(gathered).max(dim=0).values
    assert torch.equal(bits[:4].view(torch.float32), own), "bench parity: allreduced absmax != independently computed global absmax"
    # (2) codes of tensor 0 against the exact formula (scale from the independent absmax) on a sample
    x0 = xs[0]
    idx = torch.randint(0, x0.numel(), (1 << 16,), device=dev)
    s0 = torch.clamp(own[:1].cpu() / 7.0, min=1e-8).to(dev)  # CPU division: the reference's golden path
    want = torch.clamp(torch.round(x0.view(-1)[idx].float() / s0), -8, 7).to(torch.int64)
    byte = packed[0].view(-1)[idx // 2].to(torch.int64)
    got = torch.where(idx % 2 == 0, byte >> 4, byte & 15) - 8
    assert torch.equal(got, want), "bench parity spot check failed (INT4 codes)"
    # (3) N > 1: a checksum of the codes of one head, compared between the rank that owns it and a 1-GPU recomputation of the
    # same head done here on rank 0 with the global scale (bit-identical codes across world sizes)
    shard_parity = None
    if world > 1:
        csum = packed[0][:, 0].to(torch.int64).sum().reshape(1)
        allc = [torch.empty_like(csum) for _ in range(world)]
        dist.all_gather(allc, csum)
        if rank == 0:
            ok = True
            for r in range(world):
                h = head_shard_range(H, r, world)[0]
                g = torch.Generator(device=dev).manual_seed(1234 + 100000 * h)
                xr = torch.randn(B, T, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
                q = torch.clamp(torch.round(xr.float() / s0), -8, 7).to(torch.int64) + 8
                ok &= int(((q[..., 0::2] << 4) | q[..., 1::2]).sum()) == int(allc[r])
                del xr, q
            shard_parity = bool(ok)
            assert ok, "bench parity: head-sharded codes differ from the 1-GPU codes of the same head"

    per_token = run_per_token(cx, cfg, xs, packed, outs, Hl) if "per_token" in legs else None

    del xs, packed, outs, x0, tx, tp, to
    torch.cuda.empty_cache()
    collective = ("none (1 GPU)" if world == 1 else
                  "absmax words published into every rank's peer-mapped buffer by the absmax kernel's last CTA per tensor and gathered in the quantize "
                  "kernel's prologue (kvq_absmax_batched_publish / kvq_quant_batched_gather): no collective launch, 3 launches per step" if use_fused else
                  "absmax exchange through peer-mapped memory (kvq_absmax_exchange) on a side stream, two halves overlapped with the kernels" if use_xch else
                  "NCCL allreduce(MAX) of the absmax words, two halves overlapped with the kernels" + (f" (peer-memory exchange unavailable: {xch.why})" if xch is not None else ""))
    return {"value": value, "ms_step": ms_step, "clocks": clocks, "roofline": roof, "Hl": Hl, "per_token": per_token,
            "shard_parity": shard_parity, "collective": collective, "per_rank": per_rank}


def run_per_token(cx: Ctx, cfg, xs, packed, outs, Hl):
    """The cache classes' granularity (one scale per token over all (b, h, d), ops.py:323-342) at the C3 shape: the batched
    single-pass per-token quantizer (what QuantizedKVCache.init_from_prompt_past launches) for INT4 and INT8, and the
    per-token dequantize (QuantizedLayerKV.get_kv).  Reuses the C3 buffers."""
    torch, L, lib, dev = cx.torch, cx.L, cx.lib, cx.dev
    B, T, D = cfg["B"], cfg["T"], cfg["D"]
    n_t = len(xs)
    n_local = B * Hl * T * D
    sampler = cx.sampler()
    sc = [torch.empty(T, dtype=torch.bfloat16, device=dev) for _ in range(n_t)]
    sc32 = [torch.empty(T, dtype=torch.float32, device=dev) for _ in range(n_t)]
    px, pq4, ps, ps32 = L.ptr_table(xs), L.ptr_table(packed), L.ptr_table(sc), L.ptr_table(sc32)
    codes8 = [o.view(torch.int8).view(-1)[: n_local].view(B, Hl, T, D) for o in outs]  # INT8 codes live in the (idle) output buffers
    pq8 = L.ptr_table(codes8)
    sh, st = L.shape_stride(xs[0])
    stream = cx.stream()

    def quant(bits, pq):
        L.check(lib.kvq_quant_per_token_fused_batched(px.data_ptr(), L.BF16, n_t, sh, st, bits, 1e-8, pq.data_ptr(), None, 0,
                                                      ps.data_ptr(), ps32.data_ptr(), stream), "kvq_quant_per_token_fused_batched")

    out = {"config": f"C3 shape, {n_t} tensors [B={B},H={Hl} of {cfg['H']},T={T},D={D}] bf16, one scale per token; all tensors in one launch"}
    reps = 3
    ms8, _ = cx.time_launches(lambda: quant(8, pq8), reps, flush=False, warm=1)
    ms4, _ = cx.time_launches(lambda: quant(4, pq4), reps, flush=False, warm=1)
    cx.launches += 2 * reps
    out["quant_int8"] = roofline("token_fused_kernel<bf16,int8> (batched)", 3.0 * n_local * n_t, ms8, reps)
    out["quant_int4"] = roofline("token_fused_kernel<bf16,int4> (batched)", 2.5 * n_local * n_t, ms4, reps)
    # parity: token 7 of tensor 0 against the formula (scale = max|x[:, :, 7]| / 7 by CPU division)
    xt = xs[0][:, :, 7].float()
    s = torch.clamp(xt.abs().max().cpu() / 7.0, min=1e-8).to(dev)
    q = torch.clamp(torch.round(xt / s), -8, 7).to(torch.int64) + 8
    assert torch.equal(((q[..., 0::2] << 4) | q[..., 1::2]).to(torch.uint8), packed[0][:, :, 7]), "per-token INT4 codes differ from the formula"
    assert float(sc32[0][7]) == float(s.to(torch.bfloat16)), "per-token scale differs"

    shp = L.I64x4(B, Hl, T, D)

    def dequant4():
        for i in range(n_t):
            L.check(lib.kvq_dequant_int4_unpack(packed[i].data_ptr(), shp, None, sc[i].data_ptr(), L.BF16, L.PER_TOKEN, outs[i].data_ptr(),
                                                L.BF16, stream), "kvq_dequant_int4_unpack")

    msd, _ = cx.time_launches(dequant4, reps, flush=False, warm=1)
    cx.launches += n_t * reps
    out["dequant_int4"] = roofline("dequant_flat_token_kernel<bf16,int4>", 2.5 * n_local, msd / n_t, reps * n_t,
                                   note="per launch = one tensor; the group of 64 launches is timed as a whole")
    want = ((q - 8).float() * s.to(torch.bfloat16).float()).to(torch.bfloat16)
    assert torch.equal(outs[0][:, :, 7], want), "per-token dequantized values differ from the formula"
    out["clocks"] = sampler.stop()
    del codes8
    return out


# ----------------------------------------------------------------------------- C4: eviction / pooling at 128k, batched over layers


def run_c4(cx: Ctx, cfg):
    """BASELINE configs[3]: strided (w256, stride 4), block (w256, block 64, keep 8), budget (w256, 64), prefix+window (64 + 256)
    compaction and chunk-summary pooling (chunk 64, keep 256) of a Llama-3-8B-shape cache at T = 131072; all 32 layers x {K, V}
    in ONE launch per strategy (kvq_evict_gather_batched / kvq_chunk_mean_batched), KV heads sharded over the ranks, no
    collective.  L2 is flushed before every timed launch (the small outputs would otherwise be served from the 126 MB L2)."""
    torch, L, lib, dev, rank, world = cx.torch, cx.L, cx.lib, cx.dev, cx.rank, cx.world
    import kvq_b200
    from kvq_b200.cache.implementations import evict_plan, kept_token_indices
    from kvq_b200.distributed import head_shard_range

    Lyr, B, H, T, D = cfg["L"], cfg["B"], cfg["H"], cfg["T"], cfg["D"]
    lo, hi = head_shard_range(H, rank, world)
    Hl = hi - lo
    n_t = 2 * Lyr
    sampler = cx.sampler()
    xs = []
    for i in range(n_t):
        g = torch.Generator(device=dev).manual_seed(777 + i + 1000 * rank)
        xs.append(torch.randn(B, Hl, T, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16))
    px = L.ptr_table(xs)
    sh, st = L.shape_stride(xs[0])
    R = B * H * D * 2            # bytes per token over all heads (whole job)
    Rl = B * Hl * D * 2
    stream = cx.stream()
    reps = max(5, min(cx.args.steps, 10))
    out = {"config": (f"Llama-3-8B-shape synthetic KV (L={Lyr} x {{K,V}}, B={B}, H={H} KV heads, T={T}, D={D}, bf16), head-sharded over "
                      f"{world} GPU(s); one launch per strategy for all {n_t} tensors; BASELINE.json configs[3]"),
           "l2_policy": "256 MB buffer written before every timed launch"}
    strategies = [("strided", L.STRIDED, dict(window=256, stride=4)),
                  ("block", L.BLOCK, dict(window=256, block=64, keep=8)),
                  ("budget", L.BUDGET, dict(window=256, budget=64)),
                  ("prefix_window", L.PREFIX_WINDOW, dict(window=256, prefix=64))]
    for name, code, kw in strategies:
        plan = evict_plan(code, T, **kw)
        kept = int(plan.n_total)
        dst = torch.empty(n_t, B, Hl, kept, D, dtype=torch.bfloat16, device=dev)
        pd = L.ptr_table(list(dst.unbind(0)))

        def launch():
            L.check(lib.kvq_evict_gather_batched(px.data_ptr(), pd.data_ptr(), n_t, 2, sh, st, ctypes.byref(plan), None, None, stream),
                    "kvq_evict_gather_batched")

        ms, ms_min = cx.time_launches(launch, reps)
        cx.launches += reps + 3
        idx = torch.tensor(kept_token_indices(plan), device=dev)
        for i in (0, n_t - 1):
            assert torch.equal(dst[i], xs[i].index_select(2, idx)), f"c4 {name}: gathered rows differ from index_select"
        r = roofline("evict_gather_kernel (batched)", 2.0 * kept * Rl * n_t, ms, reps)
        r.update({"kept_tokens": kept, "min_launch_ms": ms_min, "gbs_whole_job": 2.0 * kept * R * n_t / (ms * 1e-3) / 1e9, "params": kw})
        out[name] = r
        del dst, pd
    # the same through the public API (trim_kv_strided over the 32-layer tuple: plan + pointer tables + allocation + one launch)
    pkv = tuple((xs[2 * i], xs[2 * i + 1]) for i in range(Lyr))
    kvq_b200.trim_kv_strided(pkv, window_size=256, stride=4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        o = kvq_b200.trim_kv_strided(pkv, window_size=256, stride=4)
    torch.cuda.synchronize()
    out["strided"]["api_wall_ms"] = (time.perf_counter() - t0) * 1e3 / 3
    cx.launches += 4
    del o
    # chunk-summary pooling
    chunk, keep = 64, 256
    old_len = T - keep
    n_chunks = (old_len + chunk - 1) // chunk
    dst = torch.empty(n_t, B, Hl, n_chunks + keep, D, dtype=torch.bfloat16, device=dev)
    pd = L.ptr_table(list(dst.unbind(0)))

    def launch_pool():
        L.check(lib.kvq_chunk_mean_batched(px.data_ptr(), pd.data_ptr(), n_t, L.BF16, sh, st, chunk, keep, stream), "kvq_chunk_mean_batched")

    ms, ms_min = cx.time_launches(launch_pool, reps)
    cx.launches += reps + 3
    want = torch.cat([xs[0][:, :, :old_len].float().view(B, Hl, n_chunks, chunk, D).mean(3).to(torch.bfloat16), xs[0][:, :, old_len:]], dim=2)
    torch.testing.assert_close(dst[0].float(), want.float(), rtol=2 ** -7, atol=1e-6)
    assert torch.equal(dst[0][:, :, n_chunks:], xs[0][:, :, old_len:])
    units = old_len + n_chunks + 2 * keep
    r = roofline("chunk_mean_kernel (batched)", float(units) * Rl * n_t, ms, reps)
    r.update({"chunks": n_chunks, "min_launch_ms": ms_min, "gbs_whole_job": float(units) * R * n_t / (ms * 1e-3) / 1e9,
              "params": {"chunk": chunk, "keep_last": keep}})
    out["chunk_pool"] = r
    out["clocks"] = sampler.stop()
    del xs, dst, pkv
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- paged store kernels


def run_paged(cx: Ctx):
    """PagedKVCache (implementations.py:10-121) kernels: whole-prompt append through the block table, one-token append,
    and get_kv's stitch.  [B=16, H=8/N, T=32768, D=128] fp16, 64-token blocks."""
    torch, L, lib, dev, rank, world = cx.torch, cx.L, cx.lib, cx.dev, cx.rank, cx.world
    import kvq_b200
    from kvq_b200.distributed import head_shard_range

    B, H, T, D, bs = 16, 8, 32768, 128, 64
    lo, hi = head_shard_range(H, rank, world)
    Hl = hi - lo
    sampler = cx.sampler()
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    k = torch.randn(B, Hl, T, D, generator=g, device=dev, dtype=torch.float32).to(torch.float16)
    v = torch.randn(B, Hl, T, D, generator=g, device=dev, dtype=torch.float32).to(torch.float16)
    cache = kvq_b200.PagedKVCache(block_size=bs, device=dev, dtype=torch.float16)
    cache.append_tokens(k, v)   # allocates the pool and the table
    kk, vv = cache.get_kv()
    assert torch.equal(kk, k) and torch.equal(vv, v), "paged: stitched cache differs from what was appended"
    del kk, vv
    cx.launches += 4
    sh, st = L.shape_stride(k)
    stream = cx.stream()
    outk = torch.empty_like(k)
    nb = B * Hl * T * D * 2
    reps = 5
    ms_a, _ = cx.time_launches(lambda: L.check(lib.kvq_paged_append(k.data_ptr(), 2, sh, st, cache._k_pool.data_ptr(), cache._table.data_ptr(),
                                                                   bs, 0, stream), "kvq_paged_append"), reps, flush=False)
    ms_g, _ = cx.time_launches(lambda: L.check(lib.kvq_paged_gather(cache._k_pool.data_ptr(), 2, L.I64x4(B, Hl, T, D), cache._table.data_ptr(), bs,
                                                                   outk.data_ptr(), stream), "kvq_paged_gather"), reps, flush=False)
    k1 = k[:, :, -1:, :]
    sh1, st1 = L.shape_stride(k1)
    ms_1, ms_1min = cx.time_launches(lambda: L.check(lib.kvq_paged_append(k1.data_ptr(), 2, sh1, st1, cache._k_pool.data_ptr(), cache._table.data_ptr(),
                                                                         bs, T - 1, stream), "kvq_paged_append"), 20, flush=True)
    cx.launches += 2 * (reps + 3) + 23
    out = {"config": f"PagedKVCache [B={B},H={Hl} of {H},T={T},D={D}] fp16, block {bs}; one K stream per launch",
           "append_prompt": roofline("paged_append_vec_kernel", 2.0 * nb, ms_a, reps),
           "stitch": roofline("paged_gather_kernel", 2.0 * nb, ms_g, reps),
           "append_one_token": {"avg_launch_ms": ms_1, "min_launch_ms": ms_1min, "bytes": 2 * B * Hl * D * 2,
                                "note": "64 KB per launch: latency-bound by construction"},
           "clocks": sampler.stop()}
    del k, v, cache, outk
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- C5: fused dequant decode attention


def run_decode(cx: Ctx):
    """Second half of the metric ("decode tok/s"): fused dequant decode attention over the C5 cache, head-sharded (no
    collective: heads are independent).  One step = the attention of all 80 layers for a batch of 8 sequences at 64k
    context; tok/s = B / step time (max over ranks).  Captured in a CUDA graph so the GPU time is not hidden behind
    Python; reported for the MHA shape (one query head per KV head) and for Llama-3-70B's real GQA (8 query heads per KV
    head), over the contiguous store and over a paged store (128-token blocks, shuffled block table)."""
    torch, dist, dev, rank, world, args = cx.torch, cx.dist, cx.dev, cx.rank, cx.world, cx.args
    import kvq_b200
    from kvq_b200.distributed import head_shard_range

    c = C5
    lo, hi = head_shard_range(c["H"], rank, world)
    Hl, B, T, D, Lyr = hi - lo, c["B"], c["T"], c["D"], c["L"]
    sampler = cx.sampler()
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    kcs = [torch.randint(-127, 128, (B, Hl, T, D), device=dev, dtype=torch.int8, generator=g) for _ in range(Lyr)]
    vcs = [torch.randint(0, 256, (B, Hl, T, D // 2), device=dev, dtype=torch.uint8, generator=g) for _ in range(Lyr)]
    ksc = torch.rand(T, device=dev, generator=g) * 0.01 + 0.01
    vsc = torch.rand(T, device=dev, generator=g) * 0.01 + 0.01
    ws = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    out = {}
    nbytes = Lyr * (B * Hl * T * (D + D // 2) + 8 * T)   # per rank per step
    # both kernel families for both shapes: "auto" is what a caller gets (CUDA-core DP4A at G = 1, INT8 tensor-core MMA at G > 1)
    variants = [("mha", 1, "auto"), ("mha_imma", 1, "imma"), ("gqa8", 8, "auto"), ("gqa8_cuda_core", 8, "cuda_core")]
    for name, G, variant in variants:
        q = torch.randn(B, Hl * G, 1, D, device=dev, dtype=torch.bfloat16, generator=g)

        # parity at the real length (T = 65536): layer 0, batch 0, every head, against fp32 dequantize + softmax on the GPU
        got = kvq_b200.decode_attention_quantized(q, kcs[0], ksc, 8, vcs[0], vsc, 4, workspace=ws, out_dtype=torch.float32, variant=variant)
        kd = kcs[0][0].float() * ksc[None, :, None]
        vb = vcs[0][0]
        vd = torch.stack([(vb >> 4).float() - 8.0, (vb & 15).float() - 8.0], dim=-1).reshape(Hl, T, D) * vsc[None, :, None]
        qf = q[0, :, 0].float().view(Hl, G, D)
        p = torch.softmax(torch.einsum("hgd,htd->hgt", qf, kd) * D ** -0.5, dim=-1)
        want = torch.einsum("hgt,htd->hgd", p, vd).reshape(Hl * G, D)
        err = float((got[0, :, 0] - want).abs().max() / want.abs().max())
        assert err < 2e-4, f"decode parity ({name}): max error {err:.3e} relative to max |out| at T={T}"
        del kd, vd, p, want, vb

        def step():
            for l in range(Lyr):
                kvq_b200.decode_attention_quantized(q, kcs[l], ksc, 8, vcs[l], vsc, 4, workspace=ws, variant=variant)

        step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(3):
            graph.replay()
        cx.sync_all()
        n = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = cx.max_over_ranks(e0.elapsed_time(e1) / n)
        cx.launches += 2 * Lyr * (n + 5)
        out[name] = {"tok_s": B / (ms * 1e-3), "ms_per_step": ms, "gbs_whole_job": nbytes * world / (ms * 1e-3) / 1e9,
                     "q_heads_per_kv_head": G, "variant": variant, "parity_max_rel_err_T65536": err,
                     "roofline": roofline(f"decode attention split kernel, G={G}, variant {variant} (+ merge)", nbytes / Lyr, ms / Lyr, n * Lyr,
                                          note="per launch pair (split + merge) = one layer; the 80-layer CUDA graph is timed as a whole")}
        del graph
    out["config"] = (f"Llama-3-70B-shape synthetic KV (L={Lyr}, {c['H']} KV heads, D={D}) batch {B} x {T} context, INT8-K / INT4-V "
                     f"per-token-scaled cache, fused dequant decode attention, head-sharded over {world} GPU(s); BASELINE.json configs[4]")
    out["launches_per_step"] = 2 * Lyr

    # ---- the same cache as a paged store: pools [n_phys, B, H, 128, row] read through a shuffled block table
    if not args.skip_paged_decode:
        del kcs, vcs
        torch.cuda.empty_cache()
        bs = 128
        nblk = T // bs
        Lp = Lyr
        kp = [torch.randint(-127, 128, (nblk, B, Hl, bs, D), device=dev, dtype=torch.int8, generator=g) for _ in range(Lp)]
        vp = [torch.randint(0, 256, (nblk, B, Hl, bs, D // 2), device=dev, dtype=torch.uint8, generator=g) for _ in range(Lp)]
        table = torch.randperm(nblk, device=dev, generator=g).to(torch.int32)
        ksp = (torch.rand(nblk, bs, device=dev, generator=g) * 0.01 + 0.01)
        vsp = (torch.rand(nblk, bs, device=dev, generator=g) * 0.01 + 0.01)
        for name, G in (("mha_paged", 1), ("gqa8_paged", 8)):
            q = torch.randn(B, Hl * G, 1, D, device=dev, dtype=torch.bfloat16, generator=g)

            def step_p():
                for l in range(Lp):
                    kvq_b200.decode_attention_paged(q, kp[l], ksp, 8, vp[l], vsp, 4, table, T, workspace=ws)

            step_p()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step_p()
            for _ in range(3):
                graph.replay()
            cx.sync_all()
            n = max(3, min(args.steps, 10))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = cx.max_over_ranks(e0.elapsed_time(e1) / n)
            cx.launches += 2 * Lp * (n + 5)
            out[name] = {"tok_s": B / (ms * 1e-3), "ms_per_step": ms, "gbs_whole_job": nbytes * world / (ms * 1e-3) / 1e9,
                         "q_heads_per_kv_head": G, "block_size": bs,
                         "roofline": roofline(f"decode attention split kernel (paged), G={G} (+ merge)", nbytes / Lyr, ms / Lp, n * Lp)}
            del graph
        del kp, vp
    else:
        del kcs, vcs
    if world > 1:
        out["append_sharded"] = run_append_sharded(cx, B, Hl, D, Lyr)
    out["clocks"] = sampler.stop()
    torch.cuda.empty_cache()
    return out


def run_append_sharded(cx: Ctx, B, Hl, D, Lyr):
    """The OTHER half of a head-sharded decode step: appending the step's token (K and V, INT8-K / INT4-V, one scale per token over
    ALL heads -- ops.py:330,:342 -- hence a max over ranks) to every layer's store through QuantizedLayerKV.append.  Two forms, same
    codes and scales: the exchange inside the quantize kernel over peer-mapped memory (one launch per tensor, what the store does
    when symmetric memory is available) and absmax kernel -> NCCL allreduce(MAX) -> quantize kernel (the fallback).  Eager calls
    (neither form is graph-captured here), device time of an 80-layer step, max over ranks."""
    torch, dev = cx.torch, cx.dev
    import kvq_b200
    from kvq_b200 import _lib as L
    from kvq_b200.distributed import shared_exchange
    g = torch.Generator(device=dev).manual_seed(99 + cx.rank)
    k1 = torch.randn(B, Hl, 1, D, device=dev, generator=g).to(torch.bfloat16)
    v1 = torch.randn(B, Hl, 1, D, device=dev, generator=g).to(torch.bfloat16)
    kp = torch.randn(B, Hl, 64, D, device=dev, generator=g).to(torch.bfloat16)
    layers = [kvq_b200.QuantizedLayerKV("mixed", dev, torch.bfloat16, group=True) for _ in range(Lyr)]
    for l in layers:
        l.append_tokens(kp, kp)
    have_x = shared_exchange(dev, True) is not None

    def step_exchange():
        for l in layers:
            l.append(k1, v1)

    def step_nccl():
        for l in layers:
            l._k.append_per_token(k1, group=True)
            l._v.append_per_token(v1, group=True)

    res = {"layers": Lyr, "tensors_per_step": 2 * Lyr, "exchange_available": have_x}
    for name, fn in (("exchange", step_exchange), ("nccl_allreduce", step_nccl)):
        if name == "exchange" and not have_x:
            continue
        for _ in range(3):
            fn()
        for l in layers:
            l._k.keep_last(64), l._v.keep_last(64)
        cx.sync_all()
        n = 10
        l0 = L.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        cx.sync_all()
        cx.launches += L.launch_count - l0
        res[name + "_ms_per_step"] = cx.max_over_ranks(e0.elapsed_time(e1) / n)
        res[name + "_launches_per_step"] = (L.launch_count - l0) // n
        for l in layers:
            l._k.keep_last(64), l._v.keep_last(64)
    if have_x:
        # the exchange form takes its step number from the device, so the whole 80-layer append can sit in a CUDA graph (as a serving
        # loop would capture it): device time without the host's 12 us per call
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            graph.capture_begin()
            step_exchange()
            graph.capture_end()
        torch.cuda.synchronize()
        for _ in range(3):
            graph.replay()
        cx.sync_all()
        n = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            graph.replay()
        e1.record()
        cx.sync_all()
        cx.launches += 2 * Lyr * (n + 4)
        res["exchange_cuda_graph_ms_per_step"] = cx.max_over_ranks(e0.elapsed_time(e1) / n)
        del graph
    return res


# ----------------------------------------------------------------------------- e2e: host buffers through the public API


def run_e2e(cx: Ctx, cfg, Hl):
    """Same metric through the public API with host buffers.  A ring of pinned host tensors feeds H2D copies on a
    copy stream, the kernels run through kvq_b200.quantize_int4_per_tensor_packed / dequantize_int4_per_tensor_packed on a
    compute stream, results (packed codes, scale, dequantized tensor) go back to pinned host memory on a third stream;
    every one of the 64 layer tensors of a step makes that round trip inside the timed region.  Also reported: the same
    pipeline shipping only codes + scales back (what a consumer that keeps the cache quantized on the host needs), and the
    PCIe ceiling measured with plain pinned copies of the same sizes running on all ranks at once."""
    torch, dev, world, args = cx.torch, cx.dev, cx.world, cx.args
    import kvq_b200

    Lyr, B, T, D = cfg["L"], cfg["B"], cfg["T"], cfg["D"]
    n_tensors = 2 * Lyr
    shape = (B, Hl, T, D)
    n = B * Hl * T * D
    ring = 4
    g = torch.Generator().manual_seed(99)
    host_in = [torch.empty(shape, dtype=torch.bfloat16).pin_memory() for _ in range(ring)]
    for h in host_in:
        h.view(torch.int16).random_(-16000, 16000, generator=g)  # finite bf16 bit patterns
    host_packed = [torch.empty((B, Hl, T, D // 2), dtype=torch.uint8).pin_memory() for _ in range(ring)]
    host_out = [torch.empty(shape, dtype=torch.bfloat16).pin_memory() for _ in range(ring)]
    host_scale = [torch.empty((), dtype=torch.bfloat16).pin_memory() for _ in range(ring)]
    dev_in = [torch.empty(shape, dtype=torch.bfloat16, device=dev) for _ in range(ring)]
    s_h2d, s_cmp, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(ring)]
    ev_done = [torch.cuda.Event() for _ in range(ring)]
    ev_free = [torch.cuda.Event() for _ in range(ring)]
    ev_host_free = [torch.cuda.Event() for _ in range(ring)]

    def one_step(ship_dequantized: bool):
        keep = []
        for i in range(n_tensors):
            r = i % ring
            with torch.cuda.stream(s_h2d):
                s_h2d.wait_event(ev_free[r])          # compute finished reading dev_in[r]
                dev_in[r].copy_(host_in[r], non_blocking=True)
                ev_in[r].record(s_h2d)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[r])
                packed, scale, last = kvq_b200.quantize_int4_per_tensor_packed(dev_in[r])
                out = kvq_b200.dequantize_int4_per_tensor_packed(packed, scale, last, torch.bfloat16)
                ev_free[r].record(s_cmp)
                ev_done[r].record(s_cmp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_done[r])
                host_packed[r].copy_(packed, non_blocking=True)
                host_scale[r].copy_(scale, non_blocking=True)
                if ship_dequantized:
                    host_out[r].copy_(out, non_blocking=True)
                packed.record_stream(s_d2h), out.record_stream(s_d2h), scale.record_stream(s_d2h)
            keep.append((packed, out))
            if len(keep) > ring:
                keep.pop(0)
        return 4 * n_tensors

    def timed(ship: bool, steps: int):
        for e in ev_free:
            e.record(s_cmp)
        cx.sync_all()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record(torch.cuda.current_stream(dev))
        s_h2d.wait_event(t0), s_cmp.wait_event(t0), s_d2h.wait_event(t0)
        wall0 = time.perf_counter()
        for _ in range(steps):
            cx.launches += one_step(ship)
        cur = torch.cuda.current_stream(dev)
        for s in (s_h2d, s_cmp, s_d2h):
            e = torch.cuda.Event()
            e.record(s)
            cur.wait_event(e)
        t1.record(cur)
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / steps, (time.perf_counter() - wall0) * 1e3 / steps

    for e in ev_free:
        e.record(s_cmp)
    one_step(True)  # warm (allocator, pinned pages)
    torch.cuda.synchronize()
    steps = max(1, min(args.steps, 2))
    ms, wall_ms = timed(True, steps)
    ms = cx.max_over_ranks(ms)
    ms_c, _ = timed(False, steps)
    ms_c = cx.max_over_ranks(ms_c)

    # PCIe ceiling: plain pinned copies of one tensor in each direction at the same time, all ranks at once
    cx.sync_all()
    reps = 6
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(torch.cuda.current_stream(dev))
    s_h2d.wait_event(t0), s_d2h.wait_event(t0)
    for i in range(reps):
        with torch.cuda.stream(s_h2d):
            dev_in[i % ring].copy_(host_in[i % ring], non_blocking=True)
        with torch.cuda.stream(s_d2h):
            host_out[i % ring].copy_(dev_in[(i + 1) % ring], non_blocking=True)
    cur = torch.cuda.current_stream(dev)
    for s in (s_h2d, s_d2h):
        e = torch.cuda.Event()
        e.record(s)
        cur.wait_event(e)
    t1.record(cur)
    torch.cuda.synchronize()
    dup_ms = cx.max_over_ranks(t0.elapsed_time(t1))
    link_gbs = reps * n * 2 / (dup_ms * 1e-3) / 1e9   # per direction, per rank, both directions busy, all ranks busy

    job_bytes = 7.0 * n * n_tensors * world
    h2d = n_tensors * n * 2 * world
    d2h = n_tensors * (n // 2 + 2 + n * 2) * world
    d2h_c = n_tensors * (n // 2 + 2) * world
    ideal_ms = max(h2d, d2h) / world / (link_gbs * 1e9) * 1e3
    ideal_ms_c = max(h2d, d2h_c) / world / (link_gbs * 1e9) * 1e3
    e2e = {"value": job_bytes / (ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": ms, "wall_ms_per_step": wall_ms, "steps": steps,
           "pcie_gbs_per_direction_per_gpu_measured": link_gbs, "pcie_bound_frac": ideal_ms / ms,
           "api": "kvq_b200.quantize_int4_per_tensor_packed + dequantize_int4_per_tensor_packed on pinned host buffers "
                  f"(ring of {ring}), 3 streams"}
    e2e_codes = {"value": job_bytes / (ms_c * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_c,
                 "ms_per_step": ms_c, "pcie_bound_frac": ideal_ms_c / ms_c,
                 "note": "same pipeline and kernels, only the packed codes and the scale go back to the host"}
    del host_in, host_packed, host_out, dev_in
    torch.cuda.empty_cache()
    return e2e, e2e_codes


# ----------------------------------------------------------------------------- C1 / C2: GPT-2 small through KVCacheBenchmarker


class StubTokenizer:
    eos_token_id = 50256

    def __init__(self, batch, length):
        import torch
        g = torch.Generator().manual_seed(0)
        self.ids = torch.randint(0, 50257, (batch, length), generator=g)

    def __call__(self, prompt, return_tensors="pt", truncation=False, max_length=None):
        ids = self.ids
        if truncation and max_length:
            ids = ids[:, :max_length]
        return type("Enc", (), {"input_ids": ids})()

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(str(int(i)) for i in ids)


def run_c2(cx: Ctx):
    """BASELINE configs[1] (C2: GPT-2 small, quant_mixed + sliding window 256, batch 16, 1024-token context, one B200) and
    configs[0] (C1: quant_int8, 1 prompt, 64 new tokens) through KVCacheBenchmarker.benchmark_method -- the reference's own
    entry point (benchmarker.py:643-832).  Random-init weights, stub tokenizer (no network).  tok/s as the reference reports it
    (new tokens of ONE sequence per second) and x batch."""
    torch, dev = cx.torch, cx.dev
    from transformers import GPT2Config, GPT2LMHeadModel
    from kvq_b200.benchmarking import KVCacheBenchmarker

    sampler = cx.sampler()
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config()).eval().half().to(dev)
    out = {"model": "GPT-2 small (124M, random init, fp16)", "new_tokens": 64}

    def run(bench, method, **kw):
        # warm-up at FULL length: the first pass over every sequence length pays CUDA's lazy kernel loading (cuBLAS / SDPA pick
        # kernels per shape), which a 4-token warm-up does not cover
        bench.benchmark_method(["x"], method=method, max_new_tokens=64, window_size=256, **kw)
        before = cx.L.launch_count
        r = bench.benchmark_method(["x"], method=method, max_new_tokens=64, window_size=256, **kw)
        cx.launches += cx.L.launch_count - before
        return r

    B2, P2 = 16, 960
    bench = KVCacheBenchmarker(model, StubTokenizer(B2, P2), device=str(dev))
    c2 = {"config": f"batch {B2}, prompt {P2} + 64 new tokens = 1024-token context; BASELINE.json configs[1]"}
    sw = {"strategy": "sliding_window", "window_size": 256}
    for key, method, kw in (("full_cache", "full_cache", {}),
                            ("sliding_window", "sliding_window", {}),
                            ("quant_mixed", "quant_mixed", {}),
                            ("quant_mixed_fused", "quant_mixed", {"fused_attention": True}),
                            ("quant_mixed_fused_window256", "quant_mixed", {"fused_attention": True, "evict": sw})):
        r = run(bench, method, **kw)
        c2[key] = {"tok_s": r["tokens_per_sec"], "tok_s_x_batch": r["tokens_per_sec"] * B2, "ms_per_step": 1e3 / r["tokens_per_sec"],
                   "est_kv_cache_mb": None if r["est_kv_cache_mb_avg"] != r["est_kv_cache_mb_avg"] else r["est_kv_cache_mb_avg"]}
    out["c2"] = c2
    bench1 = KVCacheBenchmarker(model, StubTokenizer(1, 64), device=str(dev))
    c1 = {"config": "1 prompt of 64 tokens + 64 new tokens; BASELINE.json configs[0] run on the GPU"}
    for key, method, kw in (("quant_int8", "quant_int8", {}), ("quant_int8_fused", "quant_int8", {"fused_attention": True}),
                            ("full_cache", "full_cache", {})):
        r = run(bench1, method, **kw)
        c1[key] = {"tok_s": r["tokens_per_sec"], "est_kv_cache_mb": None if r["est_kv_cache_mb_avg"] != r["est_kv_cache_mb_avg"] else r["est_kv_cache_mb_avg"]}
    out["c1"] = c1
    out["clocks"] = sampler.stop()
    del model, bench, bench1
    torch.cuda.empty_cache()
    return out


def c1_reference_cpu():
    """C1 as BASELINE.json words it: the reference's own KVCacheBenchmarker.benchmark_method("quant_int8") on the host CPU
    (unmodified reference from baseline/_ref with the two-method DynamicCache shim of SURVEY 8c).  None when absent."""
    import torch

    if load_reference_package() is None:
        return None
    try:
        from transformers import DynamicCache, GPT2Config, GPT2LMHeadModel
        if not hasattr(DynamicCache, "from_legacy_cache"):
            DynamicCache.to_legacy_cache = lambda self: tuple((l.keys, l.values) for l in self.layers)
            DynamicCache.from_legacy_cache = classmethod(lambda cls, t: cls(ddp_cache_data=t))
        RefBench = importlib.import_module("kvq_refsrc.benchmarking.benchmarker").KVCacheBenchmarker
        torch.manual_seed(0)
        model = GPT2LMHeadModel(GPT2Config()).eval()
        torch.set_num_threads(os.cpu_count() or 1)
        bench = RefBench(model, StubTokenizer(1, 64), device="cpu")
        bench.benchmark_method(["x"], method="quant_int8", max_new_tokens=2)
        r = bench.benchmark_method(["x"], method="quant_int8", max_new_tokens=32)
        return {"tok_s": r["tokens_per_sec"], "est_kv_cache_mb": r["est_kv_cache_mb_avg"], "cores": os.cpu_count(), "new_tokens": 32,
                "kind": "reference", "what": "unmodified reference KVCacheBenchmarker.benchmark_method('quant_int8'), fp32, CPU"}
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"{type(e).__name__}: {e}"}


# ----------------------------------------------------------------------------- driver


def run_gpu(args, cfg):
    cx = Ctx(args)
    torch, dist, rank, world = cx.torch, cx.dist, cx.rank, cx.world
    legs = set(LEGS) if args.workload == "all" else {"c3": {"c3", "per_token", "e2e", "cpu"}, "c4": {"c4"}, "c5": {"c5"},
                                                     "paged": {"paged"}, "c2": {"c2"}}[args.workload]
    if args.skip_e2e:
        legs.discard("e2e")
    if args.skip_cpu:
        legs.discard("cpu")
    if args.skip_decode:
        legs.discard("c5")
    if args.skip_extra:
        legs -= {"per_token", "c4", "paged", "c2"}
    c3 = run_c3(cx, cfg, legs) if "c3" in legs else None
    log("c3 done")
    c4 = run_c4(cx, C4) if "c4" in legs else None
    log("c4 done")
    decode = run_decode(cx) if "c5" in legs else None
    log("decode done")
    paged = run_paged(cx) if "paged" in legs else None
    e2e, e2e_codes = (None, None)
    if "e2e" in legs and c3 is not None:
        e2e, e2e_codes = run_e2e(cx, cfg, c3["Hl"])
    log("e2e done")
    gpt2 = run_c2(cx) if ("c2" in legs and rank == 0) else None
    cpu = cpu_torch = gpu_torch = c1_ref = None
    if rank == 0 and "cpu" in legs:
        if world == 1:
            gbs, cores, sample, _ = cpu_reference_sample(cfg, seconds=12.0)
            cpu = {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            cpu_torch = reference_torch_sample(cfg, "cpu", seconds=8.0, T_slice=1024)
            c1_ref = c1_reference_cpu() if gpt2 is not None else None
        gpu_torch = reference_torch_sample(dict(cfg, H=c3["Hl"]) if c3 else cfg, str(cx.dev), seconds=3.0, T_slice=4096)
    if rank == 0:
        if e2e is None:
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        line = {
            "metric": METRIC, "value": c3["value"] if c3 else None, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": c3["ms_step"] if c3 else None, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(cfg, world, "gpu"),
            "pct_of_8TBs_per_gpu": 100.0 * c3["value"] / world / 8000.0 if c3 else None,
            "clocks": c3["clocks"] if c3 else None, "e2e": e2e, "e2e_codes_only": e2e_codes, "gpu_launches": cx.launches,
            "roofline": c3["roofline"] if c3 else None,
            "cpu_baseline": cpu,
            "cpu_baseline_reference_torch": cpu_torch,
            "gpu_baseline_reference_torch": gpu_torch,
            "per_token": c3["per_token"] if c3 else None,
            "c4": c4, "decode": decode, "paged": paged,
            "c2": gpt2["c2"] if gpt2 else None, "c1": dict(gpt2["c1"], reference_cpu=c1_ref) if gpt2 else None,
            "gpt2_clocks": gpt2["clocks"] if gpt2 else None,
            "shard_parity": c3["shard_parity"] if c3 else None, "collective": c3["collective"] if c3 else None,
            "per_rank": c3["per_rank"] if c3 else None,
            "sm_count": cx.sm_count, "workload_flag": args.workload,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE line (the JSON): libraries that print to fd 1 (NCCL's version banner does) go to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kvq", choices=["kvq", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "c3", "c4", "c5", "paged", "c2"],
                    help="which legs to run (default: the C3 headline plus every other leg); a subset is for profiling runs")
    ap.add_argument("--layers", type=int, default=None, help="override the number of C3 layers (debug only; not a valid bench line)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: leave out the host-buffer leg")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only: leave out the CPU baseline legs")
    ap.add_argument("--skip-decode", action="store_true", help="profiling runs only: leave out the decode-attention (C5) leg")
    ap.add_argument("--skip-paged-decode", action="store_true", help="profiling runs only: leave out the paged half of the decode leg")
    ap.add_argument("--side-stream-exchange", action="store_true", help="N > 1: the stand-alone exchange kernel on a side stream, two halves (A/B against the fused publish / gather form)")
    ap.add_argument("--nccl-absmax", action="store_true", help="N > 1: exchange the absmax words with NCCL's allreduce instead of the peer-memory kernel (A/B)")
    ap.add_argument("--skip-extra", action="store_true", help="profiling runs only: leave out the per-token / C4 / paged / C2 legs")
    args = ap.parse_args()
    cfg = dict(C3)
    if args.layers:
        cfg["L"] = args.layers
    if args.impl == "reference":
        run_reference(args, cfg)
        return
    run_gpu(args, cfg)


if __name__ == "__main__":
    main()
