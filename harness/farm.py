"""Stream farm (BASELINE C5): independent command streams dealt round-robin over GPUs / ranks.

A command stream is serial through its own RDRAM, so streams are the shard boundary: rank r of W renders the streams k
with k mod W == r, and no collective touches the data path (results are per-stream RDRAM + scan-out).  Stream k runs
workload 2 + (k mod 3) of BASELINE.json with seed 1000 + k: C2 triangles, C3 fill-rate, C4 = the C2
triangle workload rendered with 4x internal upscaling and scanned out through two halving steps."""
from __future__ import annotations

import hashlib

import numpy as np

from harness import workloads as wl


def streams_for_rank(num_streams: int, rank: int, world: int) -> list[int]:
    return [k for k in range(num_streams) if k % world == rank]


KINDS = ("c2", "c3", "c4")
UPSCALING_4X = 1 << 3  # RDP_B200_FLAG_UPSCALING_4X


def stream_kind(k: int) -> str:
    return KINDS[k % 3]


def stream_flags(k: int) -> int:
    return UPSCALING_4X if stream_kind(k) == "c4" else 0


def stream_frames(k: int, frames: int, c2_triangles: int, c3_rects: int):
    rng = np.random.RandomState(1000 + k)
    if stream_kind(k) in ("c2", "c4"):
        return [wl.c2_frame(rng, c2_triangles) for _ in range(frames)]
    return [wl.c3_frame(rng, c3_rects) for _ in range(frames)]


def stream_rdram_image(k: int) -> np.ndarray:
    """Initial RDRAM of stream k: zeros + the textures its workload loads."""
    rdram = np.zeros(wl.RDRAM_SIZE, np.uint8)
    if stream_kind(k) in ("c2", "c4"):
        rdram[wl.TEX_ADDR:wl.TEX_ADDR + 8 * 2048] = wl.c2_textures()
    else:
        tex, tlut = wl.c3_assets()
        rdram[wl.TEX_ADDR:wl.TEX_ADDR + tex.size] = tex
        rdram[wl.TEX_ADDR + 0x1000:wl.TEX_ADDR + 0x1000 + tlut.size] = tlut
    return rdram


def stream_vi_registers(k: int):
    return wl.c3_vi_registers if stream_kind(k) == "c3" else wl.c2_vi_registers


def render_stream(make_processor, k: int, frames: int, c2_triangles: int, c3_rects: int) -> str:
    """Render stream k on a fresh processor (make_processor(rdram, flags)); returns a digest of its final RDRAM, hidden RDRAM and last scan-out."""
    rdram = stream_rdram_image(k)
    regs = stream_vi_registers(k)
    proc = make_processor(rdram, stream_flags(k))
    try:
        proc.begin_read_hidden_rdram()[:] = 0
        proc.end_write_hidden_rdram()
        proc.end_write_rdram()
        opts = proc.default_scanout_options()
        opts.blend_previous_frame = 1
        opts.upscale_deinterlacing = 0
        opts.downscale_steps = 2 if stream_kind(k) == "c4" else 0
        h = hashlib.sha256()
        for i, words in enumerate(stream_frames(k, frames, c2_triangles, c3_rects)):
            proc.enqueue_command_stream(words)
            for name, v in regs(i & 1).items():
                proc.set_vi_register(name, v)
            pix, w, hgt = proc.scanout_sync(opts)
            h.update(np.ascontiguousarray(pix).tobytes())
        proc.idle()
        h.update(proc.begin_read_rdram().tobytes())
        h.update((proc.begin_read_hidden_rdram() & 3).tobytes())
        return h.hexdigest()
    finally:
        proc.close()
