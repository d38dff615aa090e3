#!/usr/bin/env python
"""bench.py -- VI frames/s (and shaded Mpix/s) of the RDP hot path, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

The line's workload is BASELINE config 2 (C2): 640x480 RGBA32, ~5,000 textured Gouraud Z-buffered triangles per frame,
2-cycle combiner, bilinear TMEM sampling, fill-rect clears, 8 texture loads, SyncFull, 480i weave scan-out.

A STEP is a batch of 256 frames: 256 / S consecutive frames of each of S concurrent independent streams (S is swept,
the best S is the line's `value`).  With the driver's --steps 20 the timed region is 5,120 frames (>= 0.25 s), so the
end-to-end leg is not a pipeline fill / drain measurement.

  value    frames/s with the command streams already resident in HBM (record / replay: kernels only, no PCIe traffic)
  e2e      frames/s through the public C-ABI with HOST buffers: command words fed from pinned host memory, every frame's
           640x480 RGBA8 scan-out read back to pinned host memory inside the timed region.  The loop that makes the
           calls is native (harness/farm_driver.cpp: only include/rdp_b200.h calls), a few producer threads as an
           emulator farm would have, so no interpreter time sits between frames.
  e2e_rdram  the same with the rendered colour + depth images also returned to the caller's RDRAM every frame
           (signal_timeline / wait_for_timeline: the page-granular RDRAM mirror's device -> host side)
  roofline the tile kernel's algorithmic framebuffer bytes / its CUDA-event time vs the measured HBM peak
  configs  the other BASELINE configs, each measured the same two ways on a bounded batch: C3 fill-rate, C4-i VI-only
           sweep, C4-ii the triangle workload with 4x internal upscaling, C5 the 1,024-stream mixed farm
  cpu_baseline  the reference's own CPU renderer (Angrylion, oracle/_ref) on a bounded sample, rank 0, N=1 only

N>1: one process per GPU (torchrun), each rank renders its own independent streams (no collective on the data path: a
command stream is serial through RDRAM, streams are the natural shard); value = total frames / max time over ranks.
Each rank pins itself -- and with it its processors' command threads -- to its share of the host cores.
--impl reference times the reference's CPU renderer with all host threads on the same workload.
"""
from __future__ import annotations

# More hardware work queues than the default 8, before anything creates a CUDA context: every CommandProcessor owns
# five CUDA streams, and with 8+ processors per GPU the default makes unrelated streams share a queue and wait on
# each other (measured: +5..12 % on the end-to-end leg, nothing on the resident-input leg).
import os as _os
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from harness import farm  # noqa: E402
from harness import workloads as wl  # noqa: E402

METRIC = "vi_frames_per_s"
UNIT = "frames/s"
VI_REG = {n: i for i, n in enumerate(["Control", "Origin", "Width", "Intr", "VCurrentLine", "Timing", "VSync", "HSync", "Leap",
                                       "HStart", "VStart", "VBurst", "XScale", "YScale"])}
L2_BYTES = 126 << 20
FRAMES_PER_STEP = 256
KERNELS = {0: "front", 1: "fill_pass", 2: "span_bin", 3: "tiles", 4: "vi_filter", 5: "vi_scale", 6: "tiles_upscaled",
           7: "span_bin_upscaled", 8: "upscale_sync", 9: "vi_downscale"}
C2_WORKLOAD = ("C2: 640x480 RGBA32, %d triangle commands/frame (textured, Gouraud, Z-buffered, 2-cycle, bilinear), "
               "fill-rect clears, 8 LoadTile, SyncFull + 480i weave scan-out")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """One nvidia-smi process (rank 0 only) samples clocks / throttle reasons of the GPUs in use while the timed
    regions run.  (Round 1 ran one 20 ms poller per rank: eight of them on the 8-GPU box were part of its host-side
    scaling cliff.)"""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        self.gpus = set(int(g) for g in gpu_indices)
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                if int(f[0]) not in self.gpus:
                    continue
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "gpus_sampled": sorted(self.gpus)}


def pin_rank_to_its_cores(local_rank: int, local_world: int):
    """Every rank takes an equal, contiguous share of the cores the job may use; threads created afterwards (the
    processors' command threads, the driver's producer threads) inherit the mask."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    if local_world <= 1 or len(allowed) < 2 * local_world:
        return None
    n = len(allowed) // local_world
    mine = allowed[local_rank * n:(local_rank + 1) * n]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return mine


# ---------------------------------------------------------------------------------------------------------------------
# Workloads
# ---------------------------------------------------------------------------------------------------------------------
class Workload:
    """One BASELINE config as bench.py runs it: frames, initial RDRAM, VI registers, processor flags, scan-out options."""

    def __init__(self, kind: str, triangles: int = 5000, rects: int = 64):
        self.kind, self.triangles, self.rects = kind, triangles, rects
        self.upscale = 4 if kind == "c4ii" else 1
        self.downscale_steps = 2 if kind == "c4ii" else 0
        self.flags = {1: 0, 2: 1 << 2, 4: 1 << 3, 8: 1 << 4}[self.upscale]
        ds = min(self.downscale_steps, self.upscale.bit_length() - 1)
        self.out_dims = ((640 * self.upscale) >> ds, (480 * self.upscale) >> ds)

    def frames(self, seed: int, count: int):
        rng = np.random.RandomState(seed)
        if self.kind in ("c2", "c4ii"):
            return [wl.c2_frame(rng, self.triangles) for _ in range(count)]
        if self.kind == "c3":
            return [wl.c3_frame(rng, self.rects) for _ in range(count)]
        # c4i: nothing is rendered; a frame is a scan-out of (random) RDRAM.  One no-op command keeps the loop uniform.
        return [np.array([0, 0], np.uint32) for _ in range(count)]

    def rdram_image(self, seed: int) -> np.ndarray:
        rdram = np.zeros(wl.RDRAM_SIZE, np.uint8)
        if self.kind in ("c2", "c4ii"):
            rdram[wl.TEX_ADDR:wl.TEX_ADDR + 8 * 2048] = wl.c2_textures()
        elif self.kind == "c3":
            tex, tlut = wl.c3_assets()
            rdram[wl.TEX_ADDR:wl.TEX_ADDR + tex.size] = tex
            rdram[wl.TEX_ADDR + 0x1000:wl.TEX_ADDR + 0x1000 + tlut.size] = tlut
        else:
            rdram[:] = np.random.RandomState(seed ^ 0x51).randint(0, 256, rdram.size, dtype=np.uint8)
        return rdram

    def hidden_image(self, seed: int):
        if self.kind == "c4i":
            return np.random.RandomState(seed ^ 0x52).randint(0, 4, wl.HIDDEN_SIZE, dtype=np.uint8)
        return None

    def vi_registers(self, field: int):
        return wl.c2_vi_registers(field) if self.kind in ("c2", "c4ii") else wl.c3_vi_registers(field)

    def describe(self) -> str:
        return {
            "c2": "C2 triangles (the line's workload)",
            "c3": "C3 fill-rate: 640x480 RGBA16, %d large alpha-blended texture rectangles/frame, CI8 + TLUT, coverage AA, magic-square "
                  "dither, depth writes from primitive depth, 480i weave scan-out" % self.rects,
            "c4i": "C4-i VI-only sweep: 480i weave scan-out (AA + divot + dither filter + gamma + gamma dither, previous-field blend) of "
                   "a random 640x480 RGBA5551 framebuffer, nothing rendered",
            "c4ii": "C4-ii: the C2 triangle workload with 4x internal upscaling (native pass + 2560x1920 pass over 16 RDRAM slices), "
                    "scan-out of the upscaled domain through 2 halving steps",
        }[self.kind]


class Stream:
    """One independent command stream: its own CommandProcessor (RDRAM, CUDA streams, command thread) and frames."""

    def __init__(self, rdp, lib, torch, local_rank, workload: Workload, seed, n_frames, depth):
        self.wl = workload
        self.frames_np = workload.frames(seed, n_frames)
        # Pinned host copies of the frames (the e2e leg feeds the API from these).
        self.pinned = [torch.from_numpy(f.view(np.int32)).pin_memory() for f in self.frames_np]
        self.frames = [p.numpy().view(np.uint32) for p in self.pinned]
        self.n = n_frames
        self.rdram_t = torch.from_numpy(workload.rdram_image(seed))
        rdram = self.rdram_t.numpy()
        # Explicit RDRAM synchronisation: the stream declares what the CPU writes (nothing, after the initial upload) and asks
        # for rendered memory with signal_timeline when it wants it (the e2e_rdram leg); see include/rdp_b200.h.
        self.proc = rdp.CommandProcessor(rdram, wl.HIDDEN_SIZE, flags=workload.flags | rdp.FLAG_EXPLICIT_RDRAM_SYNC, cuda_device=local_rank)
        hidden = workload.hidden_image(seed)
        self.proc.begin_read_hidden_rdram()[:] = 0 if hidden is None else hidden
        self.proc.end_write_hidden_rdram()
        self.proc.end_write_rdram()
        self.dev = lib.rdp_b200_backend_device(self.proc._h)
        self.stream = torch.cuda.ExternalStream(self.proc.cuda_stream(), device=torch.device("cuda", local_rank))
        self.opts = self.proc.default_scanout_options()
        self.opts.blend_previous_frame = 1
        self.opts.upscale_deinterlacing = 0
        self.opts.downscale_steps = workload.downscale_steps
        up = workload.upscale
        self.outs_t = [torch.empty(640 * 576 * up * up, dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.outs = [t.numpy().view(np.uint32) for t in self.outs_t]
        self.depth = depth
        self.vi = [[(VI_REG[k], v) for k, v in workload.vi_registers(f).items()] for f in (0, 1)]
        self.vi_block = np.zeros((2, 14), np.uint32)
        for f in (0, 1):
            for r, v in self.vi[f]:
                self.vi_block[f, r] = v & 0xffffffff

    def api_frame(self, i, readback=True):
        self.proc.enqueue_command_stream(self.frames[i % self.n])
        self.proc.set_vi_registers(self.vi_block[i & 1])
        if readback:
            return self.proc.scanout_sync(self.opts, self.outs[0])
        return self.proc.scanout(self.opts)

    def close(self):
        self.proc.close()


class Bench:
    """Shared plumbing of every measured configuration on one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(self.world)))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
        self.cores = pin_rank_to_its_cores(self.local_rank, self.local_world)
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

        self.rdp = importlib.import_module("parallel-rdp_b200")
        lib = self.rdp.load_library()
        lib.rdp_b200_backend_device.restype = C.c_void_p
        lib.rdp_b200_backend_device.argtypes = [C.c_void_p]
        for name, argt, rest in (("rdpb_profile_enable", [C.c_void_p, C.c_int, C.c_uint32], C.c_int),
                                 ("rdpb_profile_read", [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p], C.c_int),
                                 ("rdpb_counters_read", [C.c_void_p, C.c_void_p, C.c_int], C.c_int),
                                 ("rdpb_record_begin", [C.c_void_p, C.c_size_t], C.c_int), ("rdpb_record_mark", [C.c_void_p], C.c_int),
                                 ("rdpb_record_end", [C.c_void_p], C.c_int), ("rdpb_record_steps", [C.c_void_p], C.c_uint32),
                                 ("rdpb_replay_step", [C.c_void_p, C.c_uint32], C.c_int), ("rdpb_sync", [C.c_void_p], C.c_int),
                                 ("rdpb_replay_interleaved", [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32], C.c_int)):
            getattr(lib, name).argtypes = argt
            getattr(lib, name).restype = rest
        self.lib = lib
        from harness import farm_native
        self.fn = farm_native
        self.farm = farm_native.load(self.rdp.PRODUCT_LIBRARY)
        self.fork = torch.cuda.Stream()
        self.join = torch.cuda.Stream()
        self.keep = []
        # Producer threads of the native e2e loop: a few, as an emulator farm would have; bounded by the rank's cores.
        budget = len(self.cores) if self.cores else (os.cpu_count() or 4)
        self.e2e_threads = max(1, min(args.e2e_threads, budget))

    # ---- plumbing ----------------------------------------------------------------------------------------------------
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def make_streams(self, workload, count, seed_of, n_first, n_other):
        return [Stream(self.rdp, self.lib, self.torch, self.local_rank, workload, seed_of(s), n_first if s == 0 else n_other,
                       self.args.e2e_depth) for s in range(count)]

    def record(self, streams, warmup):
        """Warm-up through the API, then record every distinct frame of every stream with device-resident inputs."""
        for st in streams:
            for i in range(max(warmup, 3)):
                st.api_frame(i)
            st.proc.idle()
            frame_bytes = max(f.size * 4 for f in st.frames)
            if self.lib.rdpb_record_begin(st.dev, int(st.n * (frame_bytes * 1.25 + (1 << 18)) + (8 << 20))) != 0:
                raise SystemExit("record_begin failed")
            for i in range(st.n):
                st.api_frame(i, readback=False)
                st.proc.idle()
                self.lib.rdpb_record_mark(st.dev)
            self.lib.rdpb_record_end(st.dev)
            assert self.lib.rdpb_record_steps(st.dev) >= st.n

    def replay(self, streams, first, count):
        devs = (C.c_void_p * len(streams))(*[st.dev for st in streams])
        if self.lib.rdpb_replay_interleaved(devs, len(streams), first, count) != 0:
            raise SystemExit("replay failed")

    def sync_streams(self, streams):
        for st in streams:
            self.lib.rdpb_sync(st.dev)

    def timed_replay(self, streams, frames_per_stream, warm_frames):
        """`frames_per_stream` frames of every stream, device-resident inputs.  Timed with CUDA events: the start event
        is recorded before any stream may begin, the stop event after all have finished; max over ranks."""
        torch = self.torch
        self.replay(streams, 0, warm_frames)
        self.sync_streams(streams)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.fork)
        for st in streams:
            st.stream.wait_event(e0)
        self.replay(streams, 0, frames_per_stream)
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st.stream)
            self.join.wait_event(ev)
        e1.record(self.join)
        e1.synchronize()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def farm_streams(self, streams):
        fn = self.fn
        arr = (fn.FarmStream * len(streams))()
        for i, st in enumerate(streams):
            fp, fw, op = fn.pointer_array(st.frames), fn.words_array(st.frames), fn.pointer_array(st.outs)
            s = arr[i]
            s.proc = st.proc._h
            s.frames, s.frame_words, s.num_frames, s.depth = fp, fw, st.n, st.depth
            s.vi_fields = st.vi_block.ctypes.data
            s.outs, s.out_capacity = op, st.outs[0].size
            fn.copy_options(s.opts, st.opts)
            s.expect_w, s.expect_h = st.wl.out_dims
            self.keep += [fp, fw, op]
        return arr

    def timed_e2e(self, streams, frames_per_stream, warm_frames, flags=0):
        """The same frames through the public API from host buffers (native loop), wall clock around the whole region:
        every scan-out image (and, with FARM_READ_BACK_RDRAM, every frame's rendered RDRAM) has landed in host memory
        when the clock stops.  Returns (ms, command-thread microseconds per frame)."""
        arr = self.farm_streams(streams)
        S = len(streams)
        threads = min(self.e2e_threads, S)
        for st in streams:
            st.proc.idle()
        if self.farm.farm_run(arr, S, 0, warm_frames, threads, flags) != 0:
            raise SystemExit("bench.py: scan-out failed in the e2e warm-up")
        for st in streams:
            st.proc.idle()
        self.barrier()
        host_t0 = [st.proc.host_timing() for st in streams]
        cpu0 = time.process_time()
        t0 = time.perf_counter()
        if self.farm.farm_run(arr, S, 0, frames_per_stream, threads, flags) != 0:
            raise SystemExit("bench.py: scan-out failed in the e2e leg")
        for st in streams:
            st.proc.idle()
        self.torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        cpu = time.process_time() - cpu0
        ms = self.max_over_ranks(wall * 1e3)
        host_t1 = [st.proc.host_timing() for st in streams]
        host_us = {k: round(1e6 * sum(b[k] - a[k] for a, b in zip(host_t0, host_t1)) / (frames_per_stream * S), 2) for k in host_t1[0]}
        # CPU time of the whole process (command threads, producer threads, driver threads) per frame, and cores kept busy.
        host_us["process_cpu_us"] = round(1e6 * cpu / (frames_per_stream * S), 2)
        host_us["cores_busy"] = round(cpu / wall, 2)
        self.barrier()
        return ms, host_us, threads

    def kernel_profile(self, st, frames):
        """Per-kernel CUDA-event times of one stream running alone (events on each kernel's launching stream)."""
        torch = self.torch
        self.lib.rdpb_profile_enable(st.dev, 1, 24 * frames + 64)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(st.stream)
        self.replay([st], 0, frames)
        p1.record(st.stream)
        p1.synchronize()
        profiled_ms = p0.elapsed_time(p1)
        kernel_ms = {}
        for kid, kname in KERNELS.items():
            n, total = C.c_uint32(0), C.c_double(0.0)
            self.lib.rdpb_profile_read(st.dev, kid, C.byref(n), C.byref(total))
            if n.value:
                kernel_ms[kname] = {"launches": int(n.value), "total_ms": round(total.value, 4),
                                    "us_per_frame": round(1e3 * total.value / frames, 3)}
        self.lib.rdpb_profile_enable(st.dev, 0, 0)
        return kernel_ms, profiled_ms

    def shaded_pairs(self, streams, reset=True):
        counters = (C.c_uint64 * 8)()
        total = 0
        for st in streams:
            self.lib.rdpb_counters_read(st.dev, counters, 1 if reset else 0)
            total += int(counters[1])
        return total


def parity_check(st: Stream, frames: int):
    """First `frames` frames of a stream through the API against the oracle: RDRAM, hidden RDRAM, every scan-out image."""
    from oracle import angrylion
    if not angrylion.available():
        return "not checked (oracle/_ref not built)"
    o = angrylion.Oracle(wl.RDRAM_SIZE)
    try:
        o.rdram[:] = st.wl.rdram_image(0)
        o.hidden[:] = 0
        for i in range(frames):
            o.command_stream(st.frames[i % st.n])
            for k, v in st.wl.vi_registers(i & 1).items():
                o.set_vi_register(VI_REG[k], v)
            ref_img = o.update_screen()
            pix, w, h = st.api_frame(i)
            if st.wl.upscale == 1 and (ref_img is None or ref_img.shape != pix.shape or not np.array_equal(ref_img & 0xFFFFFF, pix & 0xFFFFFF)):
                return "MISMATCH vs oracle: scan-out of frame %d" % i
        st.proc.idle()
        if not np.array_equal(st.proc.begin_read_rdram(), o.rdram):
            return "MISMATCH vs oracle: RDRAM after %d frames" % frames
        if not np.array_equal(st.proc.begin_read_hidden_rdram() & 3, o.hidden[:wl.HIDDEN_SIZE] & 3):
            return "MISMATCH vs oracle: hidden RDRAM after %d frames" % frames
    finally:
        o.close()
    if st.wl.upscale == 1:
        return "bit-exact vs oracle (%d frames: RDRAM, hidden RDRAM, every scan-out image)" % frames
    return "native RDRAM + hidden RDRAM bit-exact vs oracle after %d frames (the upscaled scan-out has no CPU oracle)" % frames


def tile_kernel_algorithmic_bytes(workload: Workload):
    """Framebuffer bytes the tile kernel must move per frame, SURVEY 8d: per touched 8x8 tile colour R+W and its hidden
    bits R+W, plus depth R+W and its hidden byte R+W.  C2's tile-kernel launch is the main pass (colour clear +
    triangles): 32-bit colour + depth over all 4800 tiles, 64*(2*4 + 2*2 + 2*2 + 2*1) B/tile.  (The Z clear is a
    fill-only pass and runs in k_fill_pass, timed separately.)  Upscaled: the same over S*S times the pixels."""
    tiles = (wl.WIDTH // 8) * (wl.HEIGHT // 8)
    bpp = 4 if workload.kind in ("c2", "c4ii") else 2
    return tiles * 64 * (2 * bpp + 2 * (bpp // 2) + 2 * 2 + 2 * 1)


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    B = Bench(args)
    torch, rank, world = B.torch, B.rank, B.world
    steps = args.steps
    stream_counts = sorted({max(1, int(x)) for x in args.streams.split(",")})
    stream_counts = [s for s in stream_counts if FRAMES_PER_STEP % s == 0] or [1]
    s_max = stream_counts[-1]

    c2 = Workload("c2", args.triangles)
    # Stream 0 alone cycles enough distinct frames that its command words exceed L2 (the "inputs larger than L2" timing
    # rule); the additional streams each add their own distinct frames on top of that.
    probe = c2.frames(2, 1)[0]
    frame_bytes = probe.size * 4
    n_full = max(8, int(np.ceil(1.12 * L2_BYTES / frame_bytes)))
    n_full = min(n_full, max(steps, 8)) if args.small else n_full
    n_extra = max(8, n_full // 4)
    # N > 1: the farm's round-robin deal (harness/farm.py) -- rank r owns the stream ids k with k mod world == r; stream k
    # is seeded 1000 + k.  N = 1 keeps the C2 definition's seed 2 for stream 0.
    def stream_seed(s):
        return 1000 + (s * world + rank) if world > 1 else 2 + s
    streams = B.make_streams(c2, s_max, stream_seed, n_full, n_extra)
    s0 = streams[0]
    prims_per_frame, _ = wl.frame_stats(s0.frames[0])

    # --- parity against the oracle (rank 0): the first frames through the API, RDRAM + every scan-out ------------------
    parity = "not checked"
    if rank == 0 and not args.no_parity:
        parity = parity_check(s0, args.parity_frames)
        if parity.startswith("MISMATCH"):
            raise SystemExit("bench.py: GPU output differs from the oracle (%s); refusing to report a number" % parity)

    B.record(streams, args.warmup)
    sampler = None
    if rank == 0:
        sampler = ClockSampler(range(world))
        sampler.start()
        # nvidia-smi needs a moment before its first sample: keep a replay load running until it reports, so the
        # clock record brackets the timed regions instead of starting after them.
        t_wait = time.perf_counter()
        while not sampler.lines and time.perf_counter() - t_wait < 3.0:
            B.replay(streams[:1], 0, 8)
            B.sync_streams(streams[:1])

    # --- value: K steps per stream count, CUDA events ---------------------------------------------------------------
    warm_frames = max(args.warmup, 3)
    sweep, best = {}, None
    for S in stream_counts:
        per_stream = steps * FRAMES_PER_STEP // S
        B.shaded_pairs(streams[:S])
        launches0 = sum(st.proc.launch_count() for st in streams[:S])
        ms = B.timed_replay(streams[:S], per_stream, warm_frames)
        launches = sum(st.proc.launch_count() for st in streams[:S]) - launches0
        fps = steps * FRAMES_PER_STEP * world / (ms / 1e3)
        sweep[S] = {"frames_per_s": round(fps, 2), "ms_per_step": round(ms / steps, 5), "launches": int(launches),
                    "shaded_pairs": B.shaded_pairs(streams[:S]), "ms": ms}
        if best is None or fps > sweep[best]["frames_per_s"]:
            best = S

    # --- roofline: one stream alone, per-kernel events on each kernel's launching stream ------------------------------
    prof_frames = min(steps * FRAMES_PER_STEP, 512)
    kernel_ms, profiled_ms = B.kernel_profile(s0, prof_frames)
    tiles = kernel_ms.get("tiles", {"launches": 0, "total_ms": 0.0})
    tiles_launches = max(tiles["launches"], 1)
    alg_per_launch = tile_kernel_algorithmic_bytes(c2)
    peak, peak_src = measured_peaks()
    achieved = (alg_per_launch * tiles_launches / 1e9) / (tiles["total_ms"] / 1e3) if tiles["total_ms"] > 0 else 0.0
    traffic = None
    try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (profiles/)
        t = json.load(open(os.path.join(ROOT, "profiles", "k_tiles_traffic.json")))
        traffic = int(t["dram_bytes_read_per_launch"]) + int(t["dram_bytes_write_per_launch"])
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "hbm", "kernel": "k_tiles", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(alg_per_launch),
                "avg_launch_us": round(1e3 * tiles["total_ms"] / tiles_launches, 3),
                "kernel_share_of_step": round(tiles["total_ms"] / profiled_ms, 4) if profiled_ms > 0 else None,
                "measured_on": "one stream running alone (per-launch CUDA events on its stream)",
                "note": "native-resolution RDP work is integer-ALU / latency bound, not HBM bound (SURVEY 8d); see DESIGN.md"}

    # --- e2e: K steps through the public API with host buffers ---------------------------------------------------------
    e2e_sweep, e2e_best = {}, None
    for S in stream_counts:
        per_stream = steps * FRAMES_PER_STEP // S
        ms, host_us, threads = B.timed_e2e(streams[:S], per_stream, warm_frames)
        fps = steps * FRAMES_PER_STEP * world / (ms / 1e3)
        e2e_sweep[S] = {"frames_per_s": round(fps, 2), "ms_per_step": round(ms / steps, 5), "command_thread_us_per_frame": host_us,
                        "producer_threads": threads}
        if e2e_best is None or fps > e2e_sweep[e2e_best]["frames_per_s"]:
            e2e_best = S
    # e2e with the rendered images returned to host RDRAM every frame, at the best stream count.
    mirror0 = [st.proc.rdram_mirror_stats() for st in streams[:e2e_best]]
    per_stream = steps * FRAMES_PER_STEP // e2e_best
    ms, _, _ = B.timed_e2e(streams[:e2e_best], per_stream, warm_frames, B.fn.FARM_READ_BACK_RDRAM)
    mirror1 = [st.proc.rdram_mirror_stats() for st in streams[:e2e_best]]
    d2h_rdram = sum(b["d2h_bytes"] - a["d2h_bytes"] for a, b in zip(mirror0, mirror1)) / ((per_stream + warm_frames) * e2e_best)
    e2e_rdram = {"value": round(steps * FRAMES_PER_STEP * world / (ms / 1e3), 2), "unit": UNIT, "streams_per_gpu": e2e_best,
                 "rdram_d2h_bytes_per_frame": int(d2h_rdram),
                 "path": "e2e + rdp_b200_signal_timeline / rdp_b200_wait_for_timeline per frame: the ranges the frame rendered to "
                         "(colour + depth image) are copied back into the caller's RDRAM on the read-back stream"}

    result = None
    if rank == 0:
        b, eb = sweep[best], e2e_sweep[e2e_best]
        h2d_frame = frame_bytes + 2 * 14 * 4
        d2h_frame = 640 * 480 * 4
        result = {
            "metric": METRIC, "value": b["frames_per_s"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": C2_WORKLOAD % prims_per_frame,
                       "streams_per_gpu": best,
                       "step": "%d frames: %d consecutive frames of each of the %d concurrent streams per GPU" % (FRAMES_PER_STEP, FRAMES_PER_STEP // best, best),
                       "frames_per_step": FRAMES_PER_STEP,
                       "distinct_frames": n_full + (best - 1) * n_extra, "command_bytes_per_frame": frame_bytes,
                       "l2": "inputs larger than L2: stream 0 alone cycles %d distinct frames = %.0f MB of command words (> 126 MB L2); "
                             "every further stream adds %d distinct frames of its own" % (n_full, n_full * frame_bytes / 1e6, n_extra),
                       "parallelism": "independent streams (BASELINE metric: 1/2/4/8 streams per B200), %d GPU(s), no collective" % world,
                       "host_cores_per_rank": len(B.cores) if B.cores else (os.cpu_count() or 0),
                       "parity": parity},
            "streams": {str(S): {"frames_per_s": v["frames_per_s"], "ms_per_step": v["ms_per_step"]} for S, v in sweep.items()},
            "mpix_shaded_per_s": round(b["shaded_pairs"] / (b["ms"] / 1e3) / 1e6 * world, 2),
            "shaded_pixels_per_frame": int(b["shaded_pairs"] / max(steps * FRAMES_PER_STEP, 1)),
            "e2e": {"value": eb["frames_per_s"], "unit": UNIT, "streams_per_gpu": e2e_best,
                    "h2d_bytes_per_step": h2d_frame * FRAMES_PER_STEP, "d2h_bytes_per_step": d2h_frame * FRAMES_PER_STEP,
                    "ms_per_step": eb["ms_per_step"], "command_thread_us_per_frame": eb["command_thread_us_per_frame"],
                    "producer_threads": eb["producer_threads"],
                    "streams": {str(S): v["frames_per_s"] for S, v in e2e_sweep.items()},
                    "command_thread_us_per_frame_by_streams": {str(S): v["command_thread_us_per_frame"] for S, v in e2e_sweep.items()},
                    "path": "rdp_b200_enqueue_command_stream_borrowed + rdp_b200_set_vi_registers + rdp_b200_scanout_async/_wait "
                            "(pinned host buffers, %d frames in flight per stream), called from harness/farm_driver.cpp; processors created with "
                            "RDP_B200_FLAG_EXPLICIT_RDRAM_SYNC (the stream writes no RDRAM after start-up, so no RDRAM bytes cross the bus; "
                            "e2e_rdram adds the per-frame read-back of the rendered images)" % args.e2e_depth},
            "e2e_rdram": e2e_rdram,
            "gpu_launches": b["launches"],
            "roofline": roofline,
            "kernel_ms": kernel_ms, "profiled_ms_per_frame": round(profiled_ms / prof_frames, 5),
        }

    for st in streams:
        st.close()
    streams = []

    # --- the other BASELINE configs, bounded batches ----------------------------------------------------------------------
    configs = {}
    wanted = [] if args.no_configs else [c.strip().lower() for c in args.configs.split(",") if c.strip()]
    for kind in ("c3", "c4i", "c4ii"):
        if kind in wanted:
            configs[kind.upper()] = run_config(B, Workload(kind, args.triangles), args)
    if "c5" in wanted:
        configs["C5"] = run_farm(B, args)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        result["configs"] = configs
        result["clocks"] = clocks
        # --- CPU baseline (rank 0, N=1 only): the reference's own CPU renderer, single thread, bounded sample -----------
        if world == 1 and not args.no_cpu_baseline:
            frames = Workload("c2", args.triangles).frames(2, 8)
            result["cpu_baseline"] = cpu_reference(frames, threads=1, budget_s=args.cpu_budget, max_frames=64)
    if world > 1:
        B.dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


def run_config(B: Bench, workload: Workload, args):
    """One of the other BASELINE configs on a bounded batch: S concurrent streams, device-resident value, e2e, kernels."""
    S = 4 if FRAMES_PER_STEP % 4 == 0 else 1
    total_frames = {"c3": 1024, "c4i": 2048, "c4ii": 256}[workload.kind]
    if args.small:
        total_frames //= 4
    per_stream = max(total_frames // S, 8)
    distinct = 8 if workload.kind != "c4i" else 2
    seed_of = (lambda s: 3000 + 16 * ord(workload.kind[-1]) + s * B.world + B.rank)
    streams = B.make_streams(workload, S, seed_of, distinct, distinct)
    out = {"workload": workload.describe(), "streams_per_gpu": S, "frames_timed": per_stream * S * B.world}
    try:
        if B.rank == 0 and not args.no_parity and workload.kind in ("c3", "c4ii"):
            out["parity"] = parity_check(streams[0], 2)
            if out["parity"].startswith("MISMATCH"):
                raise SystemExit("bench.py: %s differs from the oracle (%s)" % (workload.kind, out["parity"]))
        B.record(streams, 3)
        B.shaded_pairs(streams)
        ms = B.timed_replay(streams, per_stream, 3)
        out["frames_per_s"] = round(per_stream * S * B.world / (ms / 1e3), 2)
        shaded = B.shaded_pairs(streams)
        out["mpix_shaded_per_s"] = round(B.sum_over_ranks(shaded) / (ms / 1e3) / 1e6, 2)
        kernel_ms, profiled_ms = B.kernel_profile(streams[0], min(per_stream, 64))
        out["kernel_us_per_frame"] = {k: v["us_per_frame"] for k, v in kernel_ms.items()}
        if kernel_ms:
            out["dominant_kernel"] = max(kernel_ms, key=lambda k: kernel_ms[k]["total_ms"])
        ms, host_us, threads = B.timed_e2e(streams, per_stream, 3)
        w, h = workload.out_dims
        out["e2e"] = {"value": round(per_stream * S * B.world / (ms / 1e3), 2), "unit": UNIT,
                      "h2d_bytes_per_frame": int(max(f.size for f in streams[0].frames) * 4 + 112), "d2h_bytes_per_frame": w * h * 4,
                      "command_thread_us_per_frame": host_us, "producer_threads": threads}
        if workload.kind == "c4ii":
            # The one config where bytes dominate: the upscaled tile kernel against the HBM roofline.
            peak, peak_src = measured_peaks()
            k = kernel_ms.get("tiles_upscaled")
            if k and k["total_ms"] > 0:
                alg = tile_kernel_algorithmic_bytes(workload) * workload.upscale ** 2
                gbs = alg * k["launches"] / 1e9 / (k["total_ms"] / 1e3)
                out["roofline"] = {"bound": "hbm", "kernel": "k_tiles (upscaled 4x)", "achieved": round(gbs, 2), "peak": peak, "unit": "GB/s",
                                   "frac": round(gbs / peak, 5), "algorithmic_bytes_per_launch": int(alg),
                                   "avg_launch_us": round(1e3 * k["total_ms"] / k["launches"], 2), "peak_source": peak_src}
    finally:
        for st in streams:
            st.close()
    return out


def run_farm(B: Bench, args):
    """BASELINE C5: args.farm_streams independent streams (config 2 + k mod 3, seed 1000 + k), dealt round-robin over
    the ranks (harness/farm.py), each rank running its share through a pool of processors: a stream = reset the
    processor to the stream's initial RDRAM, then its frames with the scan-out read back every frame.  The batch is
    bounded: args.farm_frames frames per stream (the full definition is 100)."""
    torch, fn = B.torch, B.fn
    mine = farm.streams_for_rank(args.farm_streams, B.rank, B.world)
    distinct = min(2, args.farm_frames)
    c2_tris = args.triangles if not args.small else 500
    keep, jobs = [], (fn.FarmJob * len(mine))()
    images = {kind: torch.from_numpy(farm.stream_rdram_image(i)).pin_memory() for i, kind in enumerate(farm.KINDS)}
    vi = {kind: fn.vi_fields(farm.stream_vi_registers(i), VI_REG) for i, kind in enumerate(farm.KINDS)}
    for j, k in enumerate(mine):
        kind = farm.stream_kind(k)
        frames_np = farm.stream_frames(k, distinct, c2_tris, 64)
        pinned = [torch.from_numpy(f.view(np.int32)).pin_memory() for f in frames_np]
        frames = [p.numpy().view(np.uint32) for p in pinned]
        fp, fw = fn.pointer_array(frames), fn.words_array(frames)
        job = jobs[j]
        job.kind = 1 if kind == "c4" else 0
        job.rdram_image = images[kind].numpy().ctypes.data
        job.frames, job.frame_words, job.num_frames, job.frames_to_run = fp, fw, len(frames), args.farm_frames
        job.vi_fields = vi[kind].ctypes.data
        job.expect_w, job.expect_h = 640, 480
        job.downscale_steps = 2 if kind == "c4" else 0
        keep += [pinned, fp, fw]
    n_native, n_up = args.farm_pool, max(1, args.farm_pool // 4)
    slots = (fn.FarmSlot * (n_native + n_up))()
    procs = []
    for i in range(n_native + n_up):
        up = i >= n_native
        rdram_t = torch.zeros(wl.RDRAM_SIZE, dtype=torch.uint8)
        p = B.rdp.CommandProcessor(rdram_t.numpy(), wl.HIDDEN_SIZE, flags=(farm.UPSCALING_4X if up else 0) | B.rdp.FLAG_EXPLICIT_RDRAM_SYNC,
                                   cuda_device=B.local_rank)
        outs_t = [torch.empty(640 * 576, dtype=torch.int32).pin_memory() for _ in range(args.e2e_depth)]
        op = fn.pointer_array([t.numpy() for t in outs_t])
        slots[i].proc, slots[i].kind, slots[i].depth = p._h, 1 if up else 0, args.e2e_depth
        slots[i].outs, slots[i].out_capacity = op, 640 * 576
        slots[i].rdram, slots[i].rdram_bytes = rdram_t.numpy().ctypes.data, wl.RDRAM_SIZE
        procs.append(p)
        keep += [rdram_t, outs_t, op]
    done = C.c_uint64(0)
    try:
        # Warm-up: the first job of each kind (kernels loaded, buffers touched), not timed.
        warm = (fn.FarmJob * 2)()
        n_warm = 0
        for want in (0, 1):
            for j in range(len(mine)):
                if jobs[j].kind == want:
                    C.memmove(C.byref(warm[n_warm]), C.byref(jobs[j]), C.sizeof(fn.FarmJob))
                    n_warm += 1
                    break
        if n_warm and B.farm.farm_run_pool(slots, n_native + n_up, warm, n_warm, B.e2e_threads, 0, C.byref(done)) != 0:
            raise SystemExit("bench.py: farm warm-up failed")
        B.barrier()
        t0 = time.perf_counter()
        if B.farm.farm_run_pool(slots, n_native + n_up, jobs, len(mine), B.e2e_threads, 0, C.byref(done)) != 0:
            raise SystemExit("bench.py: a farm scan-out had the wrong size")
        for p in procs:
            p.idle()
        torch.cuda.synchronize()
        ms = B.max_over_ranks((time.perf_counter() - t0) * 1e3)
        frames = B.sum_over_ranks(float(done.value))
    finally:
        for p in procs:
            p.close()
    kinds = [farm.stream_kind(k) for k in range(args.farm_streams)]
    return {"workload": "C5 farm: %d independent streams, stream k = config 2 + (k mod 3) seeded 1000 + k (%d C2, %d C3, %d C4 = C2 at 4x "
                        "upscaling), dealt round-robin over %d GPU(s); per GPU a pool of %d native + %d upscaling processors"
                        % (args.farm_streams, kinds.count("c2"), kinds.count("c3"), kinds.count("c4"), B.world, n_native, n_up),
            "sample": "%d frames per stream (the definition's 100 bounded for the default run), %d frames in all" % (args.farm_frames, int(frames)),
            "scaling": "strong (the stream count is fixed, GPUs share it)",
            "e2e": {"value": round(frames / (ms / 1e3), 2), "unit": UNIT, "seconds": round(ms / 1e3, 3),
                    "path": "per stream: reset (8 MiB RDRAM image up, hidden RDRAM + TMEM cleared), then frames through "
                            "enqueue_command_stream_borrowed + set_vi_registers + scanout_async/_wait"},
            "frames_per_s": round(frames / (ms / 1e3), 2)}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference(frames, threads: int, budget_s: float, max_frames: int, steps: int | None = None, warmup: int = 0):
    """Times the reference's CPU renderer (Angrylion via oracle/_ref) on the same frames.
    threads == 1: rdp_cmd per command, exactly how the reference harness drives it.
    threads  > 1: the oracle's scanline-parallel worker pool, fed through an in-RDRAM display list."""
    from oracle import angrylion
    if not angrylion.available():
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "oracle/_ref not built"}
    parallel = threads > 1
    o = angrylion.Oracle(wl.RDRAM_SIZE, parallel=parallel, num_workers=threads if parallel else 0)
    o.rdram[:] = 0
    o.hidden[:] = 0
    o.rdram[wl.TEX_ADDR:wl.TEX_ADDR + 8 * 2048] = wl.c2_textures()

    def one(i):
        f = frames[i % len(frames)]
        if parallel:
            o.process_list(f, wl.LIST_ADDR)
        else:
            o.command_stream(f)
        for k, v in wl.c2_vi_registers(i & 1).items():
            o.set_vi_register(VI_REG[k], v)
        o.update_screen()

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    done = 0
    if steps is None:
        while done < max_frames and (time.perf_counter() - t0) < budget_s:
            one(done)
            done += 1
    else:
        for i in range(steps):
            one(i)
        done = steps
    dt = time.perf_counter() - t0
    o.close()
    return {"value": round(done / dt, 3), "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"{done} frames of the C2 stream through Angrylion ({'process_list + worker pool' if parallel else 'rdp_cmd per command'}"
                      f" + n64video_update_screen), {dt:.1f} s", "seconds": round(dt, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = Workload("c2", args.triangles).frames(2, 8)
    prims_per_frame, frame_bytes = wl.frame_stats(frames[0])
    # A step is FRAMES_PER_STEP frames, as in the product arm; here each step is a bounded sample of it so that the
    # whole run ends within a few minutes: time one frame first, then size the sample.
    probe = cpu_reference(frames, threads, budget_s=5.0, max_frames=2)
    per_frame = 1.0 / probe["value"] if probe["value"] else 1.0
    steps = args.steps
    sample = max(1, min(FRAMES_PER_STEP, int(120.0 / per_frame / max(steps, 1))))
    res = cpu_reference(frames, threads, budget_s=0, max_frames=0, steps=steps * sample, warmup=min(args.warmup, 3))
    fps = res["value"]
    out = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": steps, "warmup": args.warmup, "ms_per_step": round(1e3 * FRAMES_PER_STEP / fps, 4) if fps else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": C2_WORKLOAD % prims_per_frame,
                   "frames_per_step": FRAMES_PER_STEP,
                   "step": "%d frames; each step timed on a sample of %d of them (ms_per_step is scaled to the full step)" % (FRAMES_PER_STEP, sample),
                   "command_bytes_per_frame": frame_bytes,
                   "renderer": "Angrylion-Plus (reference's bundled CPU renderer), scanline-parallel worker pool"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "reference", "sample": res["sample"]},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20, help="timed steps; a step is %d frames" % FRAMES_PER_STEP)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--triangles", type=int, default=5000)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity-frames", type=int, default=6, help="frames of stream 0 checked against the oracle before timing")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C4 / C5 sub-results")
    ap.add_argument("--configs", default="c3,c4i,c4ii,c5", help="which of the other BASELINE configs to measure (comma list)")
    ap.add_argument("--e2e-depth", type=int, default=5, help="frames in flight per stream in the e2e leg")
    ap.add_argument("--e2e-threads", type=int, default=8, help="producer threads of the native e2e loop (per rank)")
    ap.add_argument("--streams", default="1,2,4,8,16", help="concurrent independent streams per GPU to measure (comma list, divisors "
                                                          "of %d); value/e2e report the best, the sweep is kept in the JSON line" % FRAMES_PER_STEP)
    ap.add_argument("--farm-streams", type=int, default=1024, help="C5: independent streams in the farm (all GPUs together)")
    ap.add_argument("--farm-frames", type=int, default=4, help="C5: frames per stream in the default (bounded) run")
    ap.add_argument("--farm-pool", type=int, default=16, help="C5: native processors per GPU (+ a quarter as many upscaling ones)")
    ap.add_argument("--small", action="store_true", help="fewer distinct frames (quick functional run; violates the L2 rule)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
