"""Stream farm (BASELINE C5): independent command streams dealt round-robin over GPUs / ranks.

A command stream is serial through its own RDRAM, so streams are the shard boundary: rank r of W renders the streams k
with k mod W == r, and no collective touches the data path (results are per-stream RDRAM + scan-out).  Stream k runs
workload 2 + (k mod 2) of BASELINE.json (C2 triangles, C3 fill-rate) with seed 1000 + k."""
from __future__ import annotations

import hashlib

import numpy as np

from harness import workloads as wl


def streams_for_rank(num_streams: int, rank: int, world: int) -> list[int]:
    return [k for k in range(num_streams) if k % world == rank]


def stream_kind(k: int) -> str:
    return "c2" if k % 2 == 0 else "c3"


def stream_frames(k: int, frames: int, c2_triangles: int, c3_rects: int):
    rng = np.random.RandomState(1000 + k)
    if stream_kind(k) == "c2":
        return [wl.c2_frame(rng, c2_triangles) for _ in range(frames)]
    return [wl.c3_frame(rng, c3_rects) for _ in range(frames)]


def render_stream(make_processor, k: int, frames: int, c2_triangles: int, c3_rects: int) -> str:
    """Render stream k on a fresh processor; returns a digest of its final RDRAM, hidden RDRAM and last scan-out."""
    rdram = np.zeros(wl.RDRAM_SIZE, np.uint8)
    if stream_kind(k) == "c2":
        rdram[wl.TEX_ADDR:wl.TEX_ADDR + 8 * 2048] = wl.c2_textures()
        regs = wl.c2_vi_registers
    else:
        tex, tlut = wl.c3_assets()
        rdram[wl.TEX_ADDR:wl.TEX_ADDR + tex.size] = tex
        rdram[wl.TEX_ADDR + 0x1000:wl.TEX_ADDR + 0x1000 + tlut.size] = tlut
        regs = wl.c3_vi_registers
    proc = make_processor(rdram)
    try:
        proc.begin_read_hidden_rdram()[:] = 0
        proc.end_write_hidden_rdram()
        proc.end_write_rdram()
        opts = proc.default_scanout_options()
        opts.blend_previous_frame = 1
        opts.upscale_deinterlacing = 0
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
