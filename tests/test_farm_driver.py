"""harness/farm_driver.cpp (the native loop bench.py's e2e legs and the C5 farm run in) on the host simulator: the frames
it renders through the public C-ABI equal a plain Python loop's, with and without RDRAM read-back, and the pooled farm
renders every stream exactly as a fresh processor would."""
import ctypes as C
import importlib

import numpy as np

from harness import farm, farm_native as fn, workloads as wl
from sidebyside import HOSTSIM_LIB, build_hostsim

rdp = importlib.import_module("parallel-rdp_b200")
TRIS, RECTS, FRAMES = 40, 3, 4


def _processor(k, rdram):
    p = rdp.CommandProcessor(rdram, wl.HIDDEN_SIZE, flags=farm.stream_flags(k), library=HOSTSIM_LIB)
    p.begin_read_hidden_rdram()[:] = 0
    p.end_write_hidden_rdram()
    p.end_write_rdram()
    return p


def _python_loop(k):
    rdram = farm.stream_rdram_image(k)
    p = _processor(k, rdram)
    try:
        opts = p.default_scanout_options()
        opts.blend_previous_frame = 1
        opts.upscale_deinterlacing = 0
        opts.downscale_steps = 2 if farm.stream_kind(k) == "c4" else 0
        images = []
        regs = farm.stream_vi_registers(k)
        for i, words in enumerate(farm.stream_frames(k, FRAMES, TRIS, RECTS)):
            p.enqueue_command_stream(words)
            for name, v in regs(i & 1).items():
                p.set_vi_register(name, v)
            pix, w, h = p.scanout_sync(opts)
            images.append(pix.copy())
        p.idle()
        return images, p.begin_read_rdram().copy()
    finally:
        p.close()


def test_farm_run_equals_python_loop():
    build_hostsim()
    lib = fn.load(HOSTSIM_LIB)
    ks = [0, 1, 3]  # c2, c3, c2
    keep, streams, procs, rdrams, outs_all = [], (fn.FarmStream * len(ks))(), [], [], []
    for i, k in enumerate(ks):
        rdram = farm.stream_rdram_image(k)
        p = _processor(k, rdram)
        frames = [np.ascontiguousarray(f, np.uint32) for f in farm.stream_frames(k, FRAMES, TRIS, RECTS)]
        outs = [np.zeros(640 * 576, np.uint32) for _ in range(FRAMES)]  # depth = FRAMES: every image stays around
        vi = fn.vi_fields(farm.stream_vi_registers(k), rdp.VI_REG)
        s = streams[i]
        s.proc = p._h
        fp, fw, op = fn.pointer_array(frames), fn.words_array(frames), fn.pointer_array(outs)
        s.frames, s.frame_words, s.num_frames, s.depth = fp, fw, len(frames), FRAMES
        s.vi_fields = vi.ctypes.data
        s.outs, s.out_capacity = op, outs[0].size
        o = p.default_scanout_options()
        o.blend_previous_frame = 1
        o.upscale_deinterlacing = 0
        fn.copy_options(s.opts, o)
        s.expect_w, s.expect_h = 640, 480
        keep += [frames, fp, fw, op, vi]
        procs.append(p)
        rdrams.append(rdram)
        outs_all.append(outs)
    try:
        assert lib.farm_run(streams, len(ks), 0, FRAMES, 2, fn.FARM_READ_BACK_RDRAM) == 0
        for i, k in enumerate(ks):
            images, final = _python_loop(k)
            for f in range(FRAMES):
                assert np.array_equal(outs_all[i][f][:640 * 480].reshape(480, 640), images[f]), (k, f)
            # FARM_READ_BACK_RDRAM: the rendered images are in the caller's RDRAM without any further call.
            assert np.array_equal(rdrams[i], final), k
    finally:
        for p in procs:
            p.close()


def test_pooled_farm_renders_every_stream_like_a_fresh_processor():
    build_hostsim()
    lib = fn.load(HOSTSIM_LIB)
    num_streams = 7  # kinds c2 c3 c4 c2 c3 c4 c2 over 2 native slots + 1 upscaled slot
    keep = []
    slots = (fn.FarmSlot * 3)()
    procs, slot_rdram, slot_outs = [], [], []
    for i, kind in enumerate((0, 0, 1)):
        rdram = np.zeros(wl.RDRAM_SIZE, np.uint8)
        p = rdp.CommandProcessor(rdram, wl.HIDDEN_SIZE, flags=farm.UPSCALING_4X if kind else 0, library=HOSTSIM_LIB)
        outs = [np.zeros(640 * 576, np.uint32) for _ in range(2)]
        op = fn.pointer_array(outs)
        slots[i].proc, slots[i].kind, slots[i].depth = p._h, kind, 2
        slots[i].outs, slots[i].out_capacity = op, outs[0].size
        slots[i].rdram, slots[i].rdram_bytes = rdram.ctypes.data, rdram.size
        procs.append(p)
        slot_rdram.append(rdram)
        slot_outs.append(outs)
        keep += [op]
    jobs = (fn.FarmJob * num_streams)()
    for k in range(num_streams):
        frames = [np.ascontiguousarray(f, np.uint32) for f in farm.stream_frames(k, FRAMES, TRIS, RECTS)]
        image = farm.stream_rdram_image(k)
        vi = fn.vi_fields(farm.stream_vi_registers(k), rdp.VI_REG)
        fp, fw = fn.pointer_array(frames), fn.words_array(frames)
        j = jobs[k]
        j.kind = 1 if farm.stream_kind(k) == "c4" else 0
        j.rdram_image = image.ctypes.data
        j.frames, j.frame_words, j.num_frames, j.frames_to_run = fp, fw, len(frames), FRAMES
        j.vi_fields = vi.ctypes.data
        j.expect_w, j.expect_h = 640, 480
        j.downscale_steps = 2 if farm.stream_kind(k) == "c4" else 0
        keep += [frames, image, vi, fp, fw]
    done = C.c_uint64(0)
    try:
        assert lib.farm_run_pool(slots, 3, jobs, num_streams, 2, fn.FARM_READ_BACK_RDRAM, C.byref(done)) == 0
        assert done.value == num_streams * FRAMES
        # The last job each slot ran left its final RDRAM in the slot's host range: compare with a fresh processor.  Native
        # slots took streams (0, 1, 3, 4, 6) in queue order, the upscaled slot (2, 5); which native slot ran which is
        # scheduling, so match by content
        # (any native stream may have been a slot's last).
        finals = {k: _python_loop(k)[1] for k in (5, 6, 4, 3, 1, 0)}
        assert np.array_equal(slot_rdram[2], finals[5])
        assert any(np.array_equal(slot_rdram[0], finals[k]) for k in (6, 4, 3, 1, 0))
        assert any(np.array_equal(slot_rdram[1], finals[k]) for k in (6, 4, 3, 1, 0))
    finally:
        for p in procs:
            p.close()
